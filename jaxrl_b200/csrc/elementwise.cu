// HBM-bound elementwise / row-reduction kernels around the tensor-core GEMMs:
//   RMSNorm fwd/bwd (flax nnx.RMSNorm, contract #2; reference network.py:158-172 pre-norms, :321 output norm),
//   GLU activation backward (feed_forward.py:73-77 transposed), gelu backward, inverse RoPE for dq/dk
//   (positional_embeddings.py:14-27 transposed), bias gradients (column sums), bf16 residual add.
// One warp per token row where a row reduction is needed, warp-shuffle reductions, 8-byte vector accesses.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int EW_WARPS = 8;

__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float (&f)[4]) {
  uint2 v;
  v.x = pack_bf16(f[0], f[1]);
  v.y = pack_bf16(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = v;
}

// ------------------------------------------------------------------------------------------------ RMSNorm fwd
// y = bf16( x32 * (rsqrt(mean(x32^2) + eps) * scale) ).  H % 128 == 0, H <= 1024.
template <int ITERS>
__global__ void __launch_bounds__(EW_WARPS * 32) rmsnorm_fwd_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, float eps, __nv_bfloat16* __restrict__ y,
    long long M, int H) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * EW_WARPS + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * EW_WARPS;
  griddep_launch();
  griddep_wait();
  for (; row < M; row += stride) {
    float v[ITERS][4];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      ld4(x + row * H + it * 128 + lane * 4, v[it]);
#pragma unroll
      for (int i = 0; i < 4; ++i) ss += v[it][i] * v[it][i];
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)H + eps);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const float4 s = *reinterpret_cast<const float4*>(scale + it * 128 + lane * 4);
      float o[4] = {v[it][0] * (rstd * s.x), v[it][1] * (rstd * s.y), v[it][2] * (rstd * s.z), v[it][3] * (rstd * s.w)};
      st4(y + row * H + it * 128 + lane * 4, o);
    }
  }
}

// H == 128 / 256 specialisation (every shipped config): 16 or 32 lanes per row with 16-byte accesses (two rows or one per warp
// instruction), two of them in flight per iteration - the generic kernel keeps one row per warp in flight with 8-byte accesses
// and measured 1.9 TB/s on the training tensors (17 us per 131072 rows of 128).
template <int LPR>   // lanes per row: 16 (hidden 128, two rows per warp instruction) or 32 (hidden 256)
__global__ void __launch_bounds__(EW_WARPS * 32) rmsnorm_fwdw_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, float eps, __nv_bfloat16* __restrict__ y,
    long long M) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int H = 8 * LPR, RPW = 32 / LPR;
  const int sub = lane & (LPR - 1), rsel = lane / LPR;
  float sc[8];
  {
    const float4 a = reinterpret_cast<const float4*>(scale + sub * 8)[0], b = reinterpret_cast<const float4*>(scale + sub * 8)[1];
    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
  }
  griddep_launch();
  griddep_wait();
  const long long stride = (long long)gridDim.x * EW_WARPS * 2 * RPW;
  for (long long base = ((long long)blockIdx.x * EW_WARPS + warp) * 2 * RPW; base < M; base += stride) {
    uint4 xq[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long row = base + RPW * u + rsel;
      live[u] = row < M;
      xq[u] = live[u] ? *reinterpret_cast<const uint4*>(x + row * H + sub * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t xw[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w};
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += bf16_lo(xw[e]) * bf16_lo(xw[e]) + bf16_hi(xw[e]) * bf16_hi(xw[e]);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss * (1.0f / (float)H) + eps);
      uint4 o;
      o.x = pack_bf16(bf16_lo(xw[0]) * (rstd * sc[0]), bf16_hi(xw[0]) * (rstd * sc[1]));
      o.y = pack_bf16(bf16_lo(xw[1]) * (rstd * sc[2]), bf16_hi(xw[1]) * (rstd * sc[3]));
      o.z = pack_bf16(bf16_lo(xw[2]) * (rstd * sc[4]), bf16_hi(xw[2]) * (rstd * sc[5]));
      o.w = pack_bf16(bf16_lo(xw[3]) * (rstd * sc[6]), bf16_hi(xw[3]) * (rstd * sc[7]));
      const long long row = base + RPW * u + rsel;
      if (live[u]) *reinterpret_cast<uint4*>(y + row * H + sub * 8) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------ RMSNorm bwd
// a = dy * scale ; dx = rstd * a - x * rstd^3 / H * sum(a * x) ; dscale += sum_rows dy * x * rstd
// out = bf16(dx) (+ dres in bf16 if given: the residual branch's cotangent).
template <int ITERS>
__global__ void __launch_bounds__(EW_WARPS * 32) rmsnorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
    float eps, const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dscale,
    long long M, int H) {
  __shared__ float s_ds[EW_WARPS][ITERS * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long row = (long long)blockIdx.x * EW_WARPS + warp;
  const long long stride = (long long)gridDim.x * EW_WARPS;
  float ds[ITERS][4];
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) ds[it][i] = 0.f;
  for (; row < M; row += stride) {
    float xv[ITERS][4], gv[ITERS][4];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      ld4(x + row * H + it * 128 + lane * 4, xv[it]);
      ld4(dy + row * H + it * 128 + lane * 4, gv[it]);
#pragma unroll
      for (int i = 0; i < 4; ++i) ss += xv[it][i] * xv[it][i];
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)H + eps);
    float dot = 0.f;
    float a[ITERS][4];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const float4 s = *reinterpret_cast<const float4*>(scale + it * 128 + lane * 4);
      const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[it][i] = gv[it][i] * sv[i];
        dot += a[it][i] * xv[it][i];
        ds[it][i] += gv[it][i] * xv[it][i] * rstd;
      }
    }
    dot = warp_sum(dot);
    const float c = rstd * rstd * rstd * dot / (float)H;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = rstd * a[it][i] - xv[it][i] * c;
      if (dres) {
        float r[4];
        ld4(dres + row * H + it * 128 + lane * 4, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = bf16_round(o[i]) + r[i];
      }
      st4(dx + row * H + it * 128 + lane * 4, o);
    }
  }
  if (dscale) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) s_ds[warp][it * 128 + lane * 4 + i] = ds[it][i];
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += EW_WARPS * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < EW_WARPS; ++w) t += s_ds[w][j];
      atomicAdd(dscale + j, t);
    }
  }
}

// H == 128 / 256 specialisation: 16 / 32 lanes per row with 16-byte accesses (two rows / one per warp instruction), two of them in
// flight per iteration - the generic kernel above is latency bound at this width (8-byte accesses, one row per warp).
template <int LPR>   // lanes per row: 16 (hidden 128) or 32 (hidden 256)
__global__ void __launch_bounds__(EW_WARPS * 32) rmsnorm_bwdw_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
    float eps, const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dscale,
    long long M) {
  constexpr int H = 8 * LPR, RPW = 32 / LPR;
  __shared__ float s_ds[EW_WARPS * RPW][H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (LPR - 1), rsel = lane / LPR;
  float sc[8];
  {
    const float4 a = reinterpret_cast<const float4*>(scale + sub * 8)[0], b = reinterpret_cast<const float4*>(scale + sub * 8)[1];
    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
  }
  float ds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long stride = (long long)gridDim.x * EW_WARPS * 2 * RPW;
  for (long long base = ((long long)blockIdx.x * EW_WARPS + warp) * 2 * RPW; base < M; base += stride) {
    uint4 xq[2], gq[2], rq[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long row = base + RPW * u + rsel;
      live[u] = row < M;
      xq[u] = gq[u] = rq[u] = make_uint4(0, 0, 0, 0);
      if (live[u]) {
        xq[u] = *reinterpret_cast<const uint4*>(x + row * H + sub * 8);
        gq[u] = *reinterpret_cast<const uint4*>(dy + row * H + sub * 8);
        if (dres) rq[u] = *reinterpret_cast<const uint4*>(dres + row * H + sub * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t xw[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w}, gw[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
      const uint32_t rw[4] = {rq[u].x, rq[u].y, rq[u].z, rq[u].w};
      float xv[8], gv[8], a[8];
      float ss = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xv[2 * e] = bf16_lo(xw[e]); xv[2 * e + 1] = bf16_hi(xw[e]);
        gv[2 * e] = bf16_lo(gw[e]); gv[2 * e + 1] = bf16_hi(gw[e]);
        ss += xv[2 * e] * xv[2 * e] + xv[2 * e + 1] * xv[2 * e + 1];
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss * (1.0f / (float)H) + eps);
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = gv[i] * sc[i];
        dot += a[i] * xv[i];
        ds[i] += gv[i] * xv[i] * rstd;
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const float c = rstd * rstd * rstd * dot * (1.0f / (float)H);
      uint32_t ow[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float o0 = rstd * a[2 * e] - xv[2 * e] * c, o1 = rstd * a[2 * e + 1] - xv[2 * e + 1] * c;
        if (dres) {
          const uint32_t t = pack_bf16(o0, o1);
          o0 = bf16_lo(t) + bf16_lo(rw[e]);
          o1 = bf16_hi(t) + bf16_hi(rw[e]);
        }
        ow[e] = pack_bf16(o0, o1);
      }
      if (live[u]) *reinterpret_cast<uint4*>(dx + (base + RPW * u + rsel) * H + sub * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  if (dscale) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_ds[warp * RPW + rsel][sub * 8 + i] = ds[i];
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += EW_WARPS * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < EW_WARPS * RPW; ++w) t += s_ds[w][j];
      atomicAdd(dscale + j, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GLU backward
// forward: a = bf16(gelu(u)); h = bf16(a * g).  backward: dg = bf16(dh * a); du = bf16(bf16(dh * g) * gelu'(u)).
// dug is (M, 2F): [du | dg].
__global__ void glu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ u,
                               const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dug, long long M, int F) {
  const unsigned n4 = (unsigned)(M * (F / 4));                         // host guarantees < 2^31
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned f4 = (unsigned)(F / 4);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long row = i / f4;
    const int c = (int)(i % f4) * 4;
    float dhv[4], uv[4], gv[4], du[4], dg[4];
    ld4(dh + row * F + c, dhv);
    ld4(u + row * F + c, uv);
    ld4(g + row * F + c, gv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = bf16_round(gelu_tanh(uv[k]));
      dg[k] = dhv[k] * a;
      du[k] = bf16_round(dhv[k] * gv[k]) * gelu_tanh_grad(uv[k]);
    }
    st4(dug + row * 2 * F + c, du);
    st4(dug + row * 2 * F + F + c, dg);
  }
}

// dz = bf16(dy * gelu'(z)), elementwise.
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                __nv_bfloat16* __restrict__ dz, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float a[4], b[4], o[4];
    ld4(dy + i, a);
    ld4(z + i, b);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = a[k] * gelu_tanh_grad(b[k]);
    st4(dz + i, o);
  }
}

// ------------------------------------------------------------------------------------------------ RoPE (in place)
// x (M, ld) bf16; heads [0, nheads) of 32 columns each starting at column 0 are rotated by +theta (inverse = 0)
// or -theta (inverse = 1, the transpose used for dq/dk).  theta_i = pos[row % pos_len] / timescale[i].
__global__ void rope_kernel(__nv_bfloat16* __restrict__ x, long long ld, const int32_t* __restrict__ pos, int pos_len,
                            const float* __restrict__ timescale, long long M, int nheads, int inverse) {
  // one thread per (row, head, pair-of-pairs): 8 threads per head, each 2 (i, i+16) pairs
  const long long total = M * nheads * 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int sub = (int)(i & 7);
    const long long rh = i >> 3;
    const int head = (int)(rh % nheads);
    const long long row = rh / nheads;
    const float pf = (float)pos[row % pos_len];
    __nv_bfloat16* p = x + row * ld + head * 32 + sub * 2;
    const uint32_t lo = *reinterpret_cast<const uint32_t*>(p);
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(p + 16);
    float a[2] = {bf16_lo(lo), bf16_hi(lo)}, b[2] = {bf16_lo(hi), bf16_hi(hi)}, oa[2], ob[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float sn, cs;
      rope_sincos(pf / timescale[sub * 2 + k], &sn, &cs);
      if (inverse) sn = -sn;
      oa[k] = a[k] * cs - b[k] * sn;
      ob[k] = b[k] * cs + a[k] * sn;
    }
    *reinterpret_cast<uint32_t*>(p) = pack_bf16(oa[0], oa[1]);
    *reinterpret_cast<uint32_t*>(p + 16) = pack_bf16(ob[0], ob[1]);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[N] (fp32, atomically accumulated) += sum over rows of x (M, ld) bf16 columns [0,N).
// Thread = 8 consecutive columns (one 16-byte load per row); a block owns a strip of rows, its 256 threads cover
// (256 / ceil(N/8)) row lanes; partials meet in shared memory, one atomic per column per block.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long M,
                                                     int N, float* __restrict__ out, int rows_per_block) {
  __shared__ float s_part[256 * 8];
  const int groups = (N + 7) / 8;                    // 16-byte column groups
  const int gpp = groups < 256 ? groups : 256;       // groups handled per pass
  const int RL = 256 / gpp;                          // row lanes
  const int gl = threadIdx.x % gpp, rl = threadIdx.x / gpp;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + (long long)rows_per_block);
  for (int g0 = 0; g0 < groups; g0 += gpp) {
    const int c = (g0 + gl) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rl < RL && c < N) {
      if (c + 8 <= N) {
        long long r = r0 + rl;
        for (; r + 3LL * RL < r1; r += 4LL * RL) {          // four independent 16-byte loads in flight
          uint4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(x + (r + (long long)u * RL) * ld + c);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[0] += bf16_lo(v[u].x); acc[1] += bf16_hi(v[u].x); acc[2] += bf16_lo(v[u].y); acc[3] += bf16_hi(v[u].y);
            acc[4] += bf16_lo(v[u].z); acc[5] += bf16_hi(v[u].z); acc[6] += bf16_lo(v[u].w); acc[7] += bf16_hi(v[u].w);
          }
        }
        for (; r < r1; r += RL) {
          const uint4 v = *reinterpret_cast<const uint4*>(x + r * ld + c);
          acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
          acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
        }
      } else {
        for (long long r = r0 + rl; r < r1; r += RL)
          for (int e = 0; e < 8 && c + e < N; ++e) acc[e] += __bfloat162float(x[r * ld + c + e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_part[threadIdx.x * 8 + e] = acc[e];
    __syncthreads();
    for (int j = threadIdx.x; j < gpp * 8; j += 256) {
      const int gg = j / 8, e = j % 8;
      const int col = (g0 + gg) * 8 + e;
      if (col < N) {
        float t = 0.f;
        for (int k = 0; k < RL; ++k) t += s_part[(k * gpp + gg) * 8 + e];
        atomicAdd(out + col, t);
      }
    }
    __syncthreads();
  }
}

__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float x[4], y[4], o[4];
    ld4(a + i, x);
    ld4(b + i, y);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = x[k] + y[k];
    st4(out + i, o);
  }
}

static int ew_grid(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_rmsnorm_fwd(const mpx_bf16* x, const float* scale, float eps, mpx_bf16* y, long long M, int H,
                               mpx_stream_t stream) {
  MPX_REQUIRE(x && scale && y && M > 0, "mpx_rmsnorm_fwd: null pointer or empty input");
  MPX_REQUIRE(H % 128 == 0 && H >= 128 && H <= 512, "mpx_rmsnorm_fwd: H=%d must be 128, 256, 384 or 512", H);
  const int grid = ew_grid(M, EW_WARPS);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto Y = reinterpret_cast<__nv_bfloat16*>(y);
  cudaStream_t s = (cudaStream_t)stream;
  const int fam = M <= 8192 ? 4 : 1;      // rollout-sized rows: programmatic launch like the decode kernels (gemm_tc.cu)
  if ((H == 128 || H == 256) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(scale) & 15) == 0) {
    const long long cap = (long long)num_sms() * 8;
    const int rows = H == 128 ? EW_WARPS * 4 : EW_WARPS * 2;          // rows per block and iteration
    const long long g4 = (M + rows - 1) / rows;
    const dim3 grid4((unsigned)(g4 < cap ? g4 : cap));
    if (H == 128) launch_ex(rmsnorm_fwdw_kernel<16>, grid4, dim3(EW_WARPS * 32), 0, s, 0, fam, X, scale, eps, Y, M);
    else launch_ex(rmsnorm_fwdw_kernel<32>, grid4, dim3(EW_WARPS * 32), 0, s, 0, fam, X, scale, eps, Y, M);
    return check_launch("mpx_rmsnorm_fwd");
  }
  switch (H / 128) {
    case 1: launch_ex(rmsnorm_fwd_kernel<1>, dim3(grid), dim3(EW_WARPS * 32), 0, s, 0, fam, X, scale, eps, Y, M, H); break;
    case 2: launch_ex(rmsnorm_fwd_kernel<2>, dim3(grid), dim3(EW_WARPS * 32), 0, s, 0, fam, X, scale, eps, Y, M, H); break;
    case 3: launch_ex(rmsnorm_fwd_kernel<3>, dim3(grid), dim3(EW_WARPS * 32), 0, s, 0, fam, X, scale, eps, Y, M, H); break;
    default: launch_ex(rmsnorm_fwd_kernel<4>, dim3(grid), dim3(EW_WARPS * 32), 0, s, 0, fam, X, scale, eps, Y, M, H); break;
  }
  return check_launch("mpx_rmsnorm_fwd");
}

extern "C" int mpx_rmsnorm_bwd(const mpx_bf16* dy, const mpx_bf16* x, const float* scale, float eps,
                               const mpx_bf16* dres, mpx_bf16* dx, float* dscale, long long M, int H,
                               mpx_stream_t stream) {
  MPX_REQUIRE(dy && x && scale && dx && M > 0, "mpx_rmsnorm_bwd: null pointer or empty input");
  MPX_REQUIRE(H % 128 == 0 && H >= 128 && H <= 512, "mpx_rmsnorm_bwd: H=%d must be 128, 256, 384 or 512", H);
  long long g = (M + EW_WARPS - 1) / EW_WARPS;
  const long long cap = (long long)num_sms() * 4;   // bounds the dscale atomics
  const int grid = (int)(g < cap ? g : cap);
  auto DY = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto DR = reinterpret_cast<const __nv_bfloat16*>(dres);
  auto DX = reinterpret_cast<__nv_bfloat16*>(dx);
  cudaStream_t s = (cudaStream_t)stream;
  if ((H == 128 || H == 256) && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (!dres || (reinterpret_cast<uintptr_t>(dres) & 15) == 0) &&
      (reinterpret_cast<uintptr_t>(scale) & 15) == 0) {
    const int rows = H == 128 ? EW_WARPS * 4 : EW_WARPS * 2;          // rows per block and iteration
    long long g4 = (M + rows - 1) / rows;
    const int grid4 = (int)(g4 < cap ? g4 : cap);
    if (H == 128) rmsnorm_bwdw_kernel<16><<<grid4, EW_WARPS * 32, 0, s>>>(DY, X, scale, eps, DR, DX, dscale, M);
    else rmsnorm_bwdw_kernel<32><<<grid4, EW_WARPS * 32, 0, s>>>(DY, X, scale, eps, DR, DX, dscale, M);
    return check_launch("mpx_rmsnorm_bwd");
  }
  switch (H / 128) {
    case 1: rmsnorm_bwd_kernel<1><<<grid, EW_WARPS * 32, 0, s>>>(DY, X, scale, eps, DR, DX, dscale, M, H); break;
    case 2: rmsnorm_bwd_kernel<2><<<grid, EW_WARPS * 32, 0, s>>>(DY, X, scale, eps, DR, DX, dscale, M, H); break;
    case 3: rmsnorm_bwd_kernel<3><<<grid, EW_WARPS * 32, 0, s>>>(DY, X, scale, eps, DR, DX, dscale, M, H); break;
    default: rmsnorm_bwd_kernel<4><<<grid, EW_WARPS * 32, 0, s>>>(DY, X, scale, eps, DR, DX, dscale, M, H); break;
  }
  return check_launch("mpx_rmsnorm_bwd");
}

extern "C" int mpx_glu_bwd(const mpx_bf16* dh, const mpx_bf16* u, const mpx_bf16* g, mpx_bf16* dug, long long M,
                           int F, mpx_stream_t stream) {
  MPX_REQUIRE(dh && u && g && dug && M > 0 && F > 0 && F % 4 == 0, "mpx_glu_bwd: bad arguments");
  MPX_REQUIRE(M * (F / 4) < (1LL << 31), "mpx_glu_bwd: too many elements for 32-bit indexing");
  glu_bwd_kernel<<<ew_grid(M * (F / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dh), reinterpret_cast<const __nv_bfloat16*>(u),
      reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<__nv_bfloat16*>(dug), M, F);
  return check_launch("mpx_glu_bwd");
}

extern "C" int mpx_gelu_bwd(const mpx_bf16* dy, const mpx_bf16* z, mpx_bf16* dz, long long n, mpx_stream_t stream) {
  MPX_REQUIRE(dy && z && dz && n > 0 && n % 4 == 0, "mpx_gelu_bwd: bad arguments (n %% 4 == 0 required)");
  gelu_bwd_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(z),
      reinterpret_cast<__nv_bfloat16*>(dz), n);
  return check_launch("mpx_gelu_bwd");
}

extern "C" int mpx_rope(mpx_bf16* x, long long ld, const int32_t* pos, int pos_len, const float* rope_timescale,
                        long long M, int nheads, int head_dim, int inverse, mpx_stream_t stream) {
  MPX_REQUIRE(x && pos && rope_timescale && M > 0 && nheads > 0 && pos_len > 0, "mpx_rope: bad arguments");
  MPX_REQUIRE(head_dim == 32, "mpx_rope: head_dim %d unsupported (only 32)", head_dim);
  rope_kernel<<<ew_grid(M * nheads * 8, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(x), ld, pos, pos_len, rope_timescale, M, nheads, inverse);
  return check_launch("mpx_rope");
}

extern "C" int mpx_colsum(const mpx_bf16* x, long long ld, long long M, int N, float* out, mpx_stream_t stream) {
  MPX_REQUIRE(x && out && M > 0 && N > 0, "mpx_colsum: bad arguments");
  MPX_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "mpx_colsum: x must be 16-byte aligned with ld %% 8 == 0");
  int rows_per_block = (int)((M + (long long)num_sms() * 4 - 1) / ((long long)num_sms() * 4));
  if (rows_per_block < 64) rows_per_block = 64;
  const int grid = (int)((M + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, M, N, out,
                                                        rows_per_block);
  return check_launch("mpx_colsum");
}

extern "C" int mpx_add_bf16(const mpx_bf16* a, const mpx_bf16* b, mpx_bf16* out, long long n, mpx_stream_t stream) {
  MPX_REQUIRE(a && b && out && n > 0 && n % 4 == 0, "mpx_add_bf16: bad arguments (n %% 4 == 0 required)");
  add_bf16_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
      reinterpret_cast<__nv_bfloat16*>(out), n);
  return check_launch("mpx_add_bf16");
}
