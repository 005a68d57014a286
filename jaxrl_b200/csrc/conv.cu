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
__global__ void col2im_kernel(const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ da, long long M, int IH,
                              int IW, int C, int kh, int kw, int sh, int sw, int OH, int OW) {
  const int cp = C / 8;
  const unsigned total = (unsigned)(M * IH * IW * cp);                 // host guarantees < 2^31
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = (int)(i % (unsigned)cp);
    unsigned r = i / (unsigned)cp;
    const int ix = (int)(r % (unsigned)IW);
    const int iy = (int)((r / (unsigned)IW) % (unsigned)IH);
    const long long m = r / (unsigned)(IW * IH);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ky = 0; ky < kh; ++ky) {
      const int ny = iy - ky;
      if (ny < 0 || ny % sh != 0 || ny / sh >= OH) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int nx = ix - kx;
        if (nx < 0 || nx % sw != 0 || nx / sw >= OW) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(
            dx + ((m * OH + ny / sh) * OW + nx / sw) * (long long)(kh * kw * C) + (ky * kw + kx) * C + c8 * 8);
        acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
        acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
      }
    }
    *reinterpret_cast<uint4*>(da + ((m * IH + iy) * IW + ix) * C + c8 * 8) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                   pack_bf16(acc[6], acc[7]));
  }
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
                          int sw, mpx_stream_t stream) {
  MPX_REQUIRE(dx && da && M > 0 && C % 8 == 0 && IH >= kh && IW >= kw && sh > 0 && sw > 0, "mpx_col2im: bad arguments");
  MPX_REQUIRE(M * IH * IW * (C / 8) < (1LL << 31), "mpx_col2im: too many elements for 32-bit indexing");
  const int OH = (IH - kh) / sh + 1, OW = (IW - kw) / sw + 1;
  col2im_kernel<<<cv_grid(M * IH * IW * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<__nv_bfloat16*>(da), M, IH, IW, C, kh, kw, sh, sw, OH, OW);
  return check_launch("mpx_col2im");
}
