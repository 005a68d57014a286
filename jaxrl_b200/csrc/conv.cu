// Observation-encoder support: GridCnnObsEncoder (reference mapox_trainer/model/observation.py:36-96) is
// lowered to the tcgen05 GEMM (mpx_linear / mpx_linear_wgrad) through im2col gathers:
//   conv1: one-hot(obs) (*) W1  ==  onehot_im2col(obs) [M*OH*OW, kh*kw*C (padded to 8)] . W1^T
//   conv2: im2col(a1) . W2^T ;  conv3 (3x3 VALID on a 3x3 map) is a plain Linear over the flattened map.
// The one-hot tensor of mapox.concat_one_hot (observation.py:85-87) is never materialised on its own.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

struct OneHotSpec {
  int K;            // observation components per cell
  int offs[5];      // prefix sums of the per-component one-hot sizes; offs[K] = C (total channels)
};

// out (M*OH*OW, KP) bf16, KP = round_up(kh*kw*C, 8) <= 256.  One warp per output row: lane i < taps*K loads the
// observation byte of item i = (tap, component); lane p < KP/8 owns the 16-byte piece p of the row and knows, from
// row-independent metadata computed once, which item and which value turn each of its 8 columns on.  Per row:
// one byte load, 8 shuffles + compares, one coalesced 16-byte store per piece.
__global__ void __launch_bounds__(256) onehot_im2col_kernel(const uint8_t* __restrict__ obs, __nv_bfloat16* __restrict__ out,
                                                            long long rows_total, int IH, int IW, OneHotSpec spec, int kh,
                                                            int kw, int sh, int sw, int OH, int OW, int KP) {
  const int lane = threadIdx.x & 31;
  const int C = spec.offs[spec.K];
  const int taps = kh * kw;
  const int nitems = taps * spec.K;             // <= 32 (checked by the host)
  const int pieces = KP / 8;                    // <= 32
  // item owned by this lane: cell offset inside the receptive field
  const int it_tap = lane / spec.K, it_k = lane % spec.K;
  const int it_off = ((it_tap / kw) * IW + (it_tap % kw)) * spec.K + it_k;
  // columns owned by this lane: source item lane and the value that lights the column (-1 = never)
  int src[8], want[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = lane * 8 + e;
    src[e] = 0;
    want[e] = -1;
    if (lane < pieces && col < taps * C) {
      const int tap = col / C, ch = col % C;
      int kc = 0;
      while (kc + 1 < spec.K && ch >= spec.offs[kc + 1]) ++kc;
      src[e] = tap * spec.K + kc;
      want[e] = ch - spec.offs[kc];
    }
  }
  const unsigned nrows = (unsigned)rows_total;                         // host guarantees < 2^31: 32-bit index math
  const unsigned stride = gridDim.x * (blockDim.x >> 5);
  for (unsigned row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrows; row += stride) {
    const int ox = (int)(row % (unsigned)OW);
    const int oy = (int)((row / (unsigned)OW) % (unsigned)OH);
    const long long m = row / (unsigned)(OW * OH);
    const uint8_t* cell0 = obs + ((m * IH + oy * sh) * IW + ox * sw) * spec.K;
    const int val = lane < nitems ? cell0[it_off] : 255;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int v = __shfl_sync(0xffffffffu, val, src[e]);
      if (v == want[e]) w[e >> 1] |= (e & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0
    }
    if (lane < pieces) *reinterpret_cast<uint4*>(out + (long long)row * KP + lane * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// First conv layer on the one-hot observation WITHOUT materialising it: out[row][c] = sum over (tap, component) of
// W[tap*C + offs[comp] + obs][c]  (fp32 sum of exact bf16 weights == the GEMM on the one-hot rows), then
// bf16 round, + bias (bf16), optional pre-activation save, gelu (bf16).  One thread per (row, 8 output channels).
__global__ void __launch_bounds__(256) conv1_onehot_fwd_kernel(
    const uint8_t* __restrict__ obs, const __nv_bfloat16* __restrict__ w /* (taps*C, cout) */,
    const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ z_out, __nv_bfloat16* __restrict__ a_out,
    long long rows_total, int IH, int IW, OneHotSpec spec, int kh, int kw, int sh, int sw, int OH, int OW, int cout,
    int apply_gelu) {
  extern __shared__ __align__(16) uint8_t c1_smem[];
  __nv_bfloat16* sw_ = reinterpret_cast<__nv_bfloat16*>(c1_smem);
  const int C = spec.offs[spec.K];
  const int taps = kh * kw;
  for (int i = threadIdx.x; i < taps * C * cout / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(sw_)[i] = reinterpret_cast<const uint4*>(w)[i];
  __syncthreads();
  const int cg = cout / 8;
  const unsigned total = (unsigned)(rows_total * cg);                // host guarantees < 2^31: 32-bit index math
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % (unsigned)cg);
    const unsigned row = i / (unsigned)cg;
    const int ox = (int)(row % (unsigned)OW);
    const int oy = (int)((row / (unsigned)OW) % (unsigned)OH);
    const long long m = row / (unsigned)(OW * OH);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ky = 0; ky < kh; ++ky) {
      for (int kx = 0; kx < kw; ++kx) {
        const uint8_t* cell = obs + ((m * IH + (oy * sh + ky)) * IW + (ox * sw + kx)) * spec.K;
        const int tap = ky * kw + kx;
        for (int kc = 0; kc < spec.K; ++kc) {
          const int val = cell[kc];
          if (val < spec.offs[kc + 1] - spec.offs[kc]) {
            const uint4 wv = *reinterpret_cast<const uint4*>(sw_ + (tap * C + spec.offs[kc] + val) * cout + c8 * 8);
            acc[0] += bf16_lo(wv.x); acc[1] += bf16_hi(wv.x); acc[2] += bf16_lo(wv.y); acc[3] += bf16_hi(wv.y);
            acc[4] += bf16_lo(wv.z); acc[5] += bf16_hi(wv.z); acc[6] += bf16_lo(wv.w); acc[7] += bf16_hi(wv.w);
          }
        }
      }
    }
    const uint4 bv = *reinterpret_cast<const uint4*>(bias + c8 * 8);
    const float b[8] = {bf16_lo(bv.x), bf16_hi(bv.x), bf16_lo(bv.y), bf16_hi(bv.y),
                        bf16_lo(bv.z), bf16_hi(bv.z), bf16_lo(bv.w), bf16_hi(bv.w)};
    float zf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) zf[e] = bf16_round(bf16_round(acc[e]) + b[e]);
    if (z_out)
      *reinterpret_cast<uint4*>(z_out + (long long)row * cout + c8 * 8) =
          make_uint4(pack_bf16(zf[0], zf[1]), pack_bf16(zf[2], zf[3]), pack_bf16(zf[4], zf[5]), pack_bf16(zf[6], zf[7]));
    if (apply_gelu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) zf[e] = gelu_tanh(zf[e]);
    }
    *reinterpret_cast<uint4*>(a_out + (long long)row * cout + c8 * 8) =
        make_uint4(pack_bf16(zf[0], zf[1]), pack_bf16(zf[2], zf[3]), pack_bf16(zf[4], zf[5]), pack_bf16(zf[6], zf[7]));
  }
}

// Same layer for the standard geometry (11x11 view, 3x3 stride 2 -> 5x5, 16 output channels), restructured around
// shared memory like the rollout encoder (obs_fused.cu): per tile of 32 agent-steps the observation bytes are turned
// into ONE joint code per cell (all K components; digit size_k = out of range), and each output position costs 9
// conflict-free float4 lookups per channel quad in a joint fp32 table (two bank-disjoint copies) instead of 18 bf16
// row gathers + unpacking.  Lane = (task of 8, channel quad); tasks run q-fastest so the z / a rows a warp writes are
// contiguous.
constexpr int C1J_AG = 32, C1J_MAXJ = 48, C1J_MAXK = 4;
struct Conv1JointSpec {
  int K, NJ, C;
  int sizes[C1J_MAXK], radix[C1J_MAXK], offs[C1J_MAXK + 1];
};
__global__ void __launch_bounds__(256) conv1_onehot_joint_kernel(
    const uint8_t* __restrict__ obs, const __nv_bfloat16* __restrict__ w /* (9*C, 16) */,
    const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ z_out, __nv_bfloat16* __restrict__ a_out,
    long long M, Conv1JointSpec spec, int apply_gelu) {
  extern __shared__ __align__(16) uint8_t cj_smem[];
  float* sT = reinterpret_cast<float*>(cj_smem);                        // [9][NJ][2 copies][16]
  uint8_t* sObs = cj_smem + 9 * spec.NJ * 2 * 16 * 4;                   // 32 x 121 x K bytes
  uint8_t* sCode = sObs + C1J_AG * 121 * C1J_MAXK;                      // 32 x 121 bytes
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- joint table: entry (tap, code, ch) = sum over components with digit_k < size_k of W[tap*C + off_k + digit_k][ch]
  for (int i = tid; i < 9 * spec.NJ * 16; i += 256) {
    const int ch = i & 15, code = (i >> 4) % spec.NJ, tap = (i >> 4) / spec.NJ;
    float acc = 0.f;
    int rem = code;
    for (int k = 0; k < spec.K; ++k) {
      const int digit = rem % (spec.sizes[k] + 1);
      rem /= (spec.sizes[k] + 1);
      if (digit < spec.sizes[k]) acc += __bfloat162float(w[(tap * spec.C + spec.offs[k] + digit) * 16 + ch]);
    }
    sT[((tap * spec.NJ + code) * 2) * 16 + ch] = acc;
    sT[((tap * spec.NJ + code) * 2 + 1) * 16 + ch] = acc;
  }
  const int cell_bytes = 121 * spec.K;
  const int tsub = lane >> 2, v = lane & 3;
  const float* tbase = sT + (tsub & 1) * 16 + 4 * v;
  const uint2 bv = *reinterpret_cast<const uint2*>(bias + 4 * v);
  const long long n_tiles = (M + C1J_AG - 1) / C1J_AG;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long m0 = tile * C1J_AG;
    const int nag = (int)min((long long)C1J_AG, M - m0);
    __syncthreads();                                                    // previous tile fully consumed (and table built)
    {
      const int nbytes = nag * cell_bytes;
      const uint8_t* src = obs + m0 * cell_bytes;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int nw = nbytes >> 4;
        for (int i = tid; i < nw; i += 256) reinterpret_cast<uint4*>(sObs)[i] = reinterpret_cast<const uint4*>(src)[i];
        for (int i = (nw << 4) + tid; i < nbytes; i += 256) sObs[i] = src[i];
      } else {
        for (int i = tid; i < nbytes; i += 256) sObs[i] = src[i];
      }
    }
    __syncthreads();
    {
      const int ncell = nag * 121;
      const int K = spec.K;
      const int s0 = spec.sizes[0], s1 = spec.sizes[1], s2 = spec.sizes[2], s3 = spec.sizes[3];
      const int r1 = spec.radix[1], r2 = spec.radix[2], r3 = spec.radix[3];
      for (int i4 = tid; i4 * 4 < ncell; i4 += 256) {
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
    __syncthreads();
    for (int g0 = warp; g0 < C1J_AG * 25 / 8; g0 += 16) {                // two task groups in flight per iteration
      float4 acc[2];
      int task[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g = g0 + 8 * u;
        task[u] = g * 8 + tsub;
        const int ag = task[u] / 25, q = task[u] % 25;
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < C1J_AG * 25 / 8 && ag < nag) {
          const uint8_t* c0 = sCode + ag * 121 + (q / 5) * 22 + (q % 5) * 2;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int code = c0[(tap / 3) * 11 + (tap % 3)];
            const float4 wv = *reinterpret_cast<const float4*>(tbase + (tap * spec.NJ + code) * 32);
            acc[u].x += wv.x; acc[u].y += wv.y; acc[u].z += wv.z; acc[u].w += wv.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ag = task[u] / 25;
        if (g0 + 8 * u < C1J_AG * 25 / 8 && ag < nag) {
          const uint32_t z0 = pack_bf16(acc[u].x, acc[u].y), z1 = pack_bf16(acc[u].z, acc[u].w);
          uint2 zz;
          zz.x = bf16x2_add(z0, bv.x);
          zz.y = bf16x2_add(z1, bv.y);
          const long long o = (m0 * 25 + task[u]) * 16 + 4 * v;         // row = (m0 + ag)*25 + q == m0*25 + task
          if (z_out) *reinterpret_cast<uint2*>(z_out + o) = zz;
          uint2 aa = zz;
          if (apply_gelu) {
            aa.x = gelu_tanh_bf16x2(zz.x);
            aa.y = gelu_tanh_bf16x2(zz.y);
          }
          *reinterpret_cast<uint2*>(a_out + o) = aa;
        }
      }
    }
  }
}

// NHWC im2col: a (M, IH, IW, C) bf16 -> x (M*OH*OW, kh*kw*C), C % 8 == 0; one thread per 16-byte piece.
__global__ void im2col_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ x, long long M, int IH,
                              int IW, int C, int kh, int kw, int sh, int sw, int OH, int OW) {
  const int cp = C / 8;
  const unsigned total = (unsigned)(M * OH * OW * kh * kw * cp);     // host guarantees < 2^31
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % (unsigned)cp);
    unsigned r = i / (unsigned)cp;
    const int tap = (int)(r % (unsigned)(kh * kw));
    r /= (unsigned)(kh * kw);
    const int ox = (int)(r % (unsigned)OW);
    const int oy = (int)((r / (unsigned)OW) % (unsigned)OH);
    const long long m = r / (unsigned)(OW * OH);
    const int ky = tap / kw, kx = tap % kw;
    const uint4 v = *reinterpret_cast<const uint4*>(a + ((m * IH + oy * sh + ky) * IW + ox * sw + kx) * C + c8 * 8);
    *reinterpret_cast<uint4*>(x + ((m * OH + oy) * OW + ox) * (long long)(kh * kw * C) + tap * C + c8 * 8) = v;
  }
}

// Transpose of im2col: da (M, IH, IW, C) = sum of the dx patches covering each cell (fp32 sum, bf16 store).
// z (same shape as da, nullable): also applies the activation transpose dz = bf16(da * gelu'(z)) in the same pass.
// KH/KW/SH/SW > 0: compile-time geometry (unrolled, predicated loads); 0: runtime geometry.
template <int KH, int KW, int SH, int SW>
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ z,
                                                     __nv_bfloat16* __restrict__ da, long long M, int IH, int IW, int C,
                                                     int kh_, int kw_, int sh_, int sw_, int OH, int OW) {
  const int kh = KH > 0 ? KH : kh_, kw = KW > 0 ? KW : kw_, sh = SH > 0 ? SH : sh_, sw = SW > 0 ? SW : sw_;
  const int cp = C / 8;
  const unsigned total = (unsigned)(M * IH * IW * cp);                 // host guarantees < 2^31
  const unsigned stride = gridDim.x * blockDim.x;
  const int patch = kh * kw * C;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % (unsigned)cp);
    unsigned r = i / (unsigned)cp;
    const int ix = (int)(r % (unsigned)IW);
    const int iy = (int)((r / (unsigned)IW) % (unsigned)IH);
    const long long m = r / (unsigned)(IW * IH);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __nv_bfloat16* base = dx + m * OH * OW * (long long)patch + c8 * 8;
    if (KH > 0) {
      uint4 v[KH > 0 ? KH * KW : 1];
#pragma unroll
      for (int ky = 0; ky < (KH > 0 ? KH : 1); ++ky) {
#pragma unroll
        for (int kx = 0; kx < (KW > 0 ? KW : 1); ++kx) {
          const int ny = iy - ky, nx = ix - kx;
          const bool ok = ny >= 0 && nx >= 0 && (SH == 1 || ny % sh == 0) && (SW == 1 || nx % sw == 0) && ny / sh < OH && nx / sw < OW;
          v[ky * (KW > 0 ? KW : 1) + kx] = ok ? *reinterpret_cast<const uint4*>(base + ((ny / sh) * OW + nx / sw) * patch + (ky * kw + kx) * C)
                                               : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int t = 0; t < (KH > 0 ? KH * KW : 1); ++t) {
        acc[0] += bf16_lo(v[t].x); acc[1] += bf16_hi(v[t].x); acc[2] += bf16_lo(v[t].y); acc[3] += bf16_hi(v[t].y);
        acc[4] += bf16_lo(v[t].z); acc[5] += bf16_hi(v[t].z); acc[6] += bf16_lo(v[t].w); acc[7] += bf16_hi(v[t].w);
      }
    } else {
      for (int ky = 0; ky < kh; ++ky) {
        const int ny = iy - ky;
        if (ny < 0 || ny % sh != 0 || ny / sh >= OH) continue;
        for (int kx = 0; kx < kw; ++kx) {
          const int nx = ix - kx;
          if (nx < 0 || nx % sw != 0 || nx / sw >= OW) continue;
          const uint4 v = *reinterpret_cast<const uint4*>(base + ((ny / sh) * OW + nx / sw) * patch + (ky * kw + kx) * C);
          acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
          acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
        }
      }
    }
    const long long o = ((m * IH + iy) * IW + ix) * C + c8 * 8;
    uint4 out = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                           pack_bf16(acc[6], acc[7]));
    if (z) {
      const uint4 zv = *reinterpret_cast<const uint4*>(z + o);
      out.x = pack_bf16(bf16_lo(out.x) * gelu_tanh_grad(bf16_lo(zv.x)), bf16_hi(out.x) * gelu_tanh_grad(bf16_hi(zv.x)));
      out.y = pack_bf16(bf16_lo(out.y) * gelu_tanh_grad(bf16_lo(zv.y)), bf16_hi(out.y) * gelu_tanh_grad(bf16_hi(zv.y)));
      out.z = pack_bf16(bf16_lo(out.z) * gelu_tanh_grad(bf16_lo(zv.z)), bf16_hi(out.z) * gelu_tanh_grad(bf16_hi(zv.z)));
      out.w = pack_bf16(bf16_lo(out.w) * gelu_tanh_grad(bf16_lo(zv.w)), bf16_hi(out.w) * gelu_tanh_grad(bf16_hi(zv.w)));
    }
    *reinterpret_cast<uint4*>(da + o) = out;
  }
}

// ------------------------------------------------------------------------------------------------ flattened encoder
// FlattenedObsEncoder (reference observation.py:98-145): one_hot(obs) -> Linear(C, 4) per cell -> flatten -> Linear.
// The per-cell Linear on a one-hot row is a lookup: e[j] = bf16( bf16(W[obs][j]) + bf16(b[j]) ); an out-of-range
// value is the all-zero one-hot (bias only).  out (M, KP) bf16, KP = round_up(cells*4, 8), pad columns zeroed.
__global__ void flat_embed_fwd_kernel(const uint8_t* __restrict__ obs, const float* __restrict__ w,
                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, long long M,
                                      int cells, int C, int KP) {
  const int per_row = KP / 4;                                         // 4-column groups per output row
  const unsigned total = (unsigned)(M * per_row);                    // host guarantees < 2^31
  const float b0 = bf16_round(bias[0]), b1 = bf16_round(bias[1]), b2 = bf16_round(bias[2]), b3 = bf16_round(bias[3]);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned m = i / (unsigned)per_row;
    const int cell = (int)(i % (unsigned)per_row);
    uint2 o = make_uint2(0u, 0u);
    if (cell < cells) {
      const int c = obs[(long long)m * cells + cell];
      float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
      if (c < C) {
        const float4 wv = *reinterpret_cast<const float4*>(w + c * 4);
        e0 = bf16_round(wv.x); e1 = bf16_round(wv.y); e2 = bf16_round(wv.z); e3 = bf16_round(wv.w);
      }
      o.x = pack_bf16(e0 + b0, e1 + b1);
      o.y = pack_bf16(e2 + b2, e3 + b3);
    }
    *reinterpret_cast<uint2*>(out + (long long)m * KP + cell * 4) = o;
  }
}

// Transpose: dW[c][j] += sum over (token, cell) with obs == c of dflat[token][cell*4 + j]; db[j] += sum over all.
// Per-block shared accumulators (C*4 + 4 floats), then one global atomic per entry per block.
__global__ void __launch_bounds__(256) flat_embed_bwd_kernel(const uint8_t* __restrict__ obs,
                                                             const __nv_bfloat16* __restrict__ dflat,
                                                             float* __restrict__ dw, float* __restrict__ db, long long M,
                                                             int cells, int C, int KP) {
  extern __shared__ float s_acc[];                                    // [C][4] + [4]
  for (int i = threadIdx.x; i < C * 4 + 4; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const unsigned total = (unsigned)(M * cells);
  float bl[4] = {0.f, 0.f, 0.f, 0.f};
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned m = i / (unsigned)cells;
    const int cell = (int)(i % (unsigned)cells);
    const uint2 v = *reinterpret_cast<const uint2*>(dflat + (long long)m * KP + cell * 4);
    const float d[4] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y)};
    const int c = obs[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bl[j] += d[j];
      if (c < C) atomicAdd(&s_acc[c * 4 + j], d[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float t = warp_sum(bl[j]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[C * 4 + j], t);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 4; i += blockDim.x)
    if (s_acc[i] != 0.f) atomicAdd(dw + i, s_acc[i]);
  if (threadIdx.x < 4) atomicAdd(db + threadIdx.x, s_acc[C * 4 + threadIdx.x]);
}

static int cv_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_obs_onehot_im2col(const uint8_t* obs, mpx_bf16* out, long long M, int IH, int IW, int K,
                                     const int* onehot_sizes, int kh, int kw, int sh, int sw, int KP,
                                     mpx_stream_t stream) {
  MPX_REQUIRE(obs && out && onehot_sizes && M > 0, "mpx_obs_onehot_im2col: bad arguments");
  MPX_REQUIRE(K >= 1 && K <= 4, "mpx_obs_onehot_im2col: K=%d components unsupported (1..4)", K);
  OneHotSpec spec;
  spec.K = K;
  spec.offs[0] = 0;
  for (int i = 0; i < K; ++i) spec.offs[i + 1] = spec.offs[i] + onehot_sizes[i];
  const int C = spec.offs[K];
  MPX_REQUIRE(KP % 8 == 0 && KP >= kh * kw * C, "mpx_obs_onehot_im2col: KP=%d must be a multiple of 8 >= %d", KP, kh * kw * C);
  MPX_REQUIRE(IH >= kh && IW >= kw && sh > 0 && sw > 0, "mpx_obs_onehot_im2col: bad geometry");
  const int OH = (IH - kh) / sh + 1, OW = (IW - kw) / sw + 1;
  const long long rows = M * OH * OW;
  MPX_REQUIRE(kh * kw * K <= 32 && KP <= 256, "mpx_obs_onehot_im2col: receptive field too large (kh*kw*K=%d, KP=%d)",
              kh * kw * K, KP);
  MPX_REQUIRE(rows < (1LL << 31), "mpx_obs_onehot_im2col: too many rows for 32-bit indexing");
  long long g = (rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  onehot_im2col_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(obs, reinterpret_cast<__nv_bfloat16*>(out), rows, IH, IW,
                                                               spec, kh, kw, sh, sw, OH, OW, KP);
  return check_launch("mpx_obs_onehot_im2col");
}

extern "C" int mpx_conv1_onehot_fwd(const uint8_t* obs, const mpx_bf16* w, const mpx_bf16* bias, mpx_bf16* z_out,
                                    mpx_bf16* a_out, long long M, int IH, int IW, int K, const int* onehot_sizes, int kh,
                                    int kw, int sh, int sw, int cout, int apply_gelu, mpx_stream_t stream) {
  MPX_REQUIRE(obs && w && bias && a_out && onehot_sizes && M > 0, "mpx_conv1_onehot_fwd: bad arguments");
  MPX_REQUIRE(K >= 1 && K <= 4 && cout % 8 == 0, "mpx_conv1_onehot_fwd: K=%d / cout=%d unsupported", K, cout);
  MPX_REQUIRE(IH >= kh && IW >= kw && sh > 0 && sw > 0, "mpx_conv1_onehot_fwd: bad geometry");
  OneHotSpec spec;
  spec.K = K;
  spec.offs[0] = 0;
  for (int i = 0; i < K; ++i) spec.offs[i + 1] = spec.offs[i] + onehot_sizes[i];
  const int C = spec.offs[K];
  const size_t smem = (size_t)kh * kw * C * cout * 2;
  MPX_REQUIRE(smem <= 48 * 1024 && (kh * kw * C * cout) % 8 == 0, "mpx_conv1_onehot_fwd: kernel too large for shared memory");
  const int OH = (IH - kh) / sh + 1, OW = (IW - kw) / sw + 1;
  const long long rows = M * OH * OW;
  MPX_REQUIRE(rows * (cout / 8) < (1LL << 31), "mpx_conv1_onehot_fwd: too many rows for 32-bit indexing");
  {
    // standard geometry: shared-memory joint-table kernel
    int nj = 1;
    for (int i = 0; i < K; ++i) nj *= onehot_sizes[i] + 1;
    if (IH == 11 && IW == 11 && kh == 3 && kw == 3 && sh == 2 && sw == 2 && cout == 16 && nj <= C1J_MAXJ &&
        (reinterpret_cast<uintptr_t>(a_out) & 7) == 0 && (!z_out || (reinterpret_cast<uintptr_t>(z_out) & 7) == 0) &&
        (reinterpret_cast<uintptr_t>(bias) & 7) == 0) {
      Conv1JointSpec js = {};
      js.K = K; js.NJ = nj; js.C = C;
      int radix = 1;
      for (int i = 0; i < C1J_MAXK; ++i) {
        js.sizes[i] = i < K ? onehot_sizes[i] : 0;
        js.radix[i] = i < K ? radix : 0;
        js.offs[i] = spec.offs[i < K ? i : K];
        if (i < K) radix *= onehot_sizes[i] + 1;
      }
      js.offs[C1J_MAXK] = C;
      const int jsmem = 9 * nj * 2 * 16 * 4 + C1J_AG * 121 * C1J_MAXK + C1J_AG * 121 + 16;
      static bool attr_set = false;
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv1_onehot_joint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) {
          set_error("cudaFuncSetAttribute(conv1_onehot_joint): %s", cudaGetErrorString(e));
          return MPX_ERR_CUDA;
        }
        attr_set = true;
      }
      const long long tiles = (M + C1J_AG - 1) / C1J_AG;
      long long grid = (long long)num_sms() * 3;
      if (grid > tiles) grid = tiles;
      conv1_onehot_joint_kernel<<<(int)grid, 256, jsmem, (cudaStream_t)stream>>>(
          obs, reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<const __nv_bfloat16*>(bias),
          reinterpret_cast<__nv_bfloat16*>(z_out), reinterpret_cast<__nv_bfloat16*>(a_out), M, js, apply_gelu);
      return check_launch("mpx_conv1_onehot_fwd(joint)");
    }
  }
  conv1_onehot_fwd_kernel<<<cv_grid(rows * (cout / 8)), 256, smem, (cudaStream_t)stream>>>(
      obs, reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<const __nv_bfloat16*>(bias),
      reinterpret_cast<__nv_bfloat16*>(z_out), reinterpret_cast<__nv_bfloat16*>(a_out), rows, IH, IW, spec, kh, kw, sh, sw,
      OH, OW, cout, apply_gelu);
  return check_launch("mpx_conv1_onehot_fwd");
}

extern "C" int mpx_im2col(const mpx_bf16* a, mpx_bf16* x, long long M, int IH, int IW, int C, int kh, int kw, int sh,
                          int sw, mpx_stream_t stream) {
  MPX_REQUIRE(a && x && M > 0 && C % 8 == 0 && IH >= kh && IW >= kw && sh > 0 && sw > 0, "mpx_im2col: bad arguments");
  const int OH = (IH - kh) / sh + 1, OW = (IW - kw) / sw + 1;
  MPX_REQUIRE(M * OH * OW * kh * kw * (C / 8) < (1LL << 31), "mpx_im2col: too many elements for 32-bit indexing");
  im2col_kernel<<<cv_grid(M * OH * OW * kh * kw * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<__nv_bfloat16*>(x), M, IH, IW, C, kh, kw, sh, sw, OH, OW);
  return check_launch("mpx_im2col");
}

extern "C" int mpx_col2im(const mpx_bf16* dx, mpx_bf16* da, long long M, int IH, int IW, int C, int kh, int kw, int sh,
                          int sw, const mpx_bf16* z, mpx_stream_t stream) {
  MPX_REQUIRE(dx && da && M > 0 && C % 8 == 0 && IH >= kh && IW >= kw && sh > 0 && sw > 0, "mpx_col2im: bad arguments");
  MPX_REQUIRE(M * IH * IW * (C / 8) < (1LL << 31), "mpx_col2im: too many elements for 32-bit indexing");
  const int OH = (IH - kh) / sh + 1, OW = (IW - kw) / sw + 1;
  const int grid = cv_grid(M * IH * IW * (C / 8));
  if (kh == 3 && kw == 3 && sh == 1 && sw == 1)
    col2im_kernel<3, 3, 1, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(z),
        reinterpret_cast<__nv_bfloat16*>(da), M, IH, IW, C, kh, kw, sh, sw, OH, OW);
  else
    col2im_kernel<0, 0, 0, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(z),
        reinterpret_cast<__nv_bfloat16*>(da), M, IH, IW, C, kh, kw, sh, sw, OH, OW);
  return check_launch("mpx_col2im");
}

extern "C" int mpx_flat_embed_fwd(const uint8_t* obs, const float* w, const float* bias, mpx_bf16* out, long long M,
                                  int cells, int C, int KP, mpx_stream_t stream) {
  MPX_REQUIRE(obs && w && bias && out && M > 0 && cells > 0 && C > 0, "mpx_flat_embed_fwd: bad arguments");
  MPX_REQUIRE(KP % 8 == 0 && KP >= cells * 4, "mpx_flat_embed_fwd: KP=%d must be a multiple of 8 >= 4*cells", KP);
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "mpx_flat_embed_fwd: embedding kernel must be 16-byte aligned");
  MPX_REQUIRE(M * (KP / 4) < (1LL << 31), "mpx_flat_embed_fwd: too many elements for 32-bit indexing");
  flat_embed_fwd_kernel<<<cv_grid(M * (KP / 4)), 256, 0, (cudaStream_t)stream>>>(
      obs, w, bias, reinterpret_cast<__nv_bfloat16*>(out), M, cells, C, KP);
  return check_launch("mpx_flat_embed_fwd");
}

extern "C" int mpx_flat_embed_bwd(const uint8_t* obs, const mpx_bf16* dflat, float* dw, float* dbias, long long M, int cells,
                                  int C, int KP, mpx_stream_t stream) {
  MPX_REQUIRE(obs && dflat && dw && dbias && M > 0 && cells > 0 && C > 0, "mpx_flat_embed_bwd: bad arguments");
  MPX_REQUIRE(KP % 8 == 0 && KP >= cells * 4 && C <= 1024, "mpx_flat_embed_bwd: bad KP / C");
  MPX_REQUIRE(M * cells < (1LL << 31), "mpx_flat_embed_bwd: too many elements for 32-bit indexing");
  long long g = (M * cells + 255) / 256;
  const long long cap = (long long)num_sms() * 4;                    // bounds the global atomics
  if (g > cap) g = cap;
  flat_embed_bwd_kernel<<<(int)g, 256, (C * 4 + 4) * sizeof(float), (cudaStream_t)stream>>>(
      obs, reinterpret_cast<const __nv_bfloat16*>(dflat), dw, dbias, M, cells, C, KP);
  return check_launch("mpx_flat_embed_bwd");
}
