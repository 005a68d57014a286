// HL-Gauss value head: expected value of the categorical and the soft-label cross-entropy loss (+ gradient).
// Restates HlGaussValue.get_value / get_loss (reference mapox_trainer/values.py:43-63):
//   value = sum_j softmax(logits)_j * centers_j
//   t = clip(target, min, max); cdf_j = Phi((s_j - t) / sigma) on the NL+1 bin edges s;
//   p_j = (cdf_{j+1} - cdf_j) / (cdf_NL - cdf_0); loss = mean_n( -sum_j p_j * log_softmax(logits)_j )
//   d loss / d logits_j = (softmax_j * sum_k p_k - p_j) / n
// All fp32 (contract #9).  HBM-bound: 2*NL*4 + 4 B per token; one warp per token row, warp-shuffle reductions.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int HLG_MAX_PER_LANE = 4;   // NL <= 128
constexpr int HLG_WARPS = 8;

// jax.scipy.special.ndtr (behind jax.scipy.stats.norm.cdf, values.py:53) in fp32.
__device__ __forceinline__ float ndtr_f32(float x) {
  const float kHalfSqrt2 = 0.70710678118654752440f;
  float w = x * kHalfSqrt2;
  float z = fabsf(w);
  float y;
  if (z < kHalfSqrt2) y = 1.0f + erff(w);
  else if (w > 0.0f) y = 2.0f - erfcf(z);
  else y = erfcf(z);
  return 0.5f * y;
}

__global__ void __launch_bounds__(HLG_WARPS * 32) hlgauss_value_kernel(
    const float* __restrict__ logits, long long n, int NL, const float* __restrict__ centers,
    float* __restrict__ value) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * HLG_WARPS + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * HLG_WARPS;
  for (; row < n; row += stride) {
    const float* x = logits + row * NL;
    float v[HLG_MAX_PER_LANE];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < HLG_MAX_PER_LANE; ++i) {
      int j = lane + 32 * i;
      v[i] = j < NL ? x[j] : -INFINITY;
      m = fmaxf(m, v[i]);
    }
    m = warp_max(m);
    float s = 0.f, sc = 0.f;
#pragma unroll
    for (int i = 0; i < HLG_MAX_PER_LANE; ++i) {
      int j = lane + 32 * i;
      if (j < NL) {
        float e = expf(v[i] - m);
        s += e;
        sc += e * centers[j];
      }
    }
    s = warp_sum(s);
    sc = warp_sum(sc);
    if (lane == 0) value[row] = sc / s;
  }
}

__global__ void __launch_bounds__(HLG_WARPS * 32) hlgauss_loss_kernel(
    const float* __restrict__ logits, const float* __restrict__ targets, long long n, int NL,
    const float* __restrict__ supports, float vmin, float vmax, float sigma, float gscale,
    float* __restrict__ dlogits, __nv_bfloat16* __restrict__ dl_bf16, int ld_bf16, double* __restrict__ partials) {
  __shared__ float s_cdf[HLG_WARPS][HLG_MAX_PER_LANE * 32 + 1];
  __shared__ double s_part[HLG_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  long long row = (long long)blockIdx.x * HLG_WARPS + warp;
  const long long stride = (long long)gridDim.x * HLG_WARPS;
  double acc = 0.0;
  for (; row < n; row += stride) {
    const float* x = logits + row * NL;
    float t = fminf(fmaxf(targets[row], vmin), vmax);
    for (int j = lane; j <= NL; j += 32) s_cdf[warp][j] = ndtr_f32((supports[j] - t) / sigma);
    __syncwarp();
    const float z = s_cdf[warp][NL] - s_cdf[warp][0];
    float v[HLG_MAX_PER_LANE], p[HLG_MAX_PER_LANE];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < HLG_MAX_PER_LANE; ++i) {
      int j = lane + 32 * i;
      v[i] = j < NL ? x[j] : -INFINITY;
      p[i] = j < NL ? (s_cdf[warp][j + 1] - s_cdf[warp][j]) / z : 0.f;
      m = fmaxf(m, v[i]);
    }
    __syncwarp();
    m = warp_max(m);
    float s = 0.f, psum = 0.f;
#pragma unroll
    for (int i = 0; i < HLG_MAX_PER_LANE; ++i) {
      int j = lane + 32 * i;
      if (j < NL) {
        s += expf(v[i] - m);
        psum += p[i];
      }
    }
    s = warp_sum(s);
    psum = warp_sum(psum);
    const float lse = logf(s);
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < HLG_MAX_PER_LANE; ++i) {
      int j = lane + 32 * i;
      if (j < NL) {
        float lsm = (v[i] - m) - lse;
        l -= p[i] * lsm;
        const float gr = gscale * (expf(lsm) * psum - p[i]);
        if (dlogits) dlogits[row * NL + j] = gr;
        if (dl_bf16) dl_bf16[row * ld_bf16 + j] = __float2bfloat16_rn(gr);
      } else if (dl_bf16 && j < ld_bf16) {
        dl_bf16[row * ld_bf16 + j] = __float2bfloat16_rn(0.f);
      }
    }
    l = warp_sum(l);
    acc += (double)l;
  }
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < HLG_WARPS; ++w) a += s_part[w];
    partials[blockIdx.x] = a;
  }
}

// NL < 64: lane l owns bins 2l and 2l+1 (bin edges 2l, 2l+1 and, from lane l+1, 2l+2) - no shared memory, half the
// instructions of the generic kernel above for the reference's 51 bins.
__global__ void __launch_bounds__(HLG_WARPS * 32) hlgauss_loss2_kernel(
    const float* __restrict__ logits, const float* __restrict__ targets, long long n, int NL,
    const float* __restrict__ supports, float vmin, float vmax, float sigma, float gscale,
    float* __restrict__ dlogits, __nv_bfloat16* __restrict__ dl_bf16, int ld_bf16, double* __restrict__ partials) {
  __shared__ double s_part[HLG_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  long long row = (long long)blockIdx.x * HLG_WARPS + warp;
  const long long stride = (long long)gridDim.x * HLG_WARPS;
  const int j0 = 2 * lane, j1 = 2 * lane + 1;
  const float e0 = j0 <= NL ? supports[j0] : 0.f, e1 = j1 <= NL ? supports[j1] : 0.f;
  const bool b0 = j0 < NL, b1 = j1 < NL;
  double acc = 0.0;
  for (; row < n; row += stride) {
    const float* x = logits + row * NL;
    const float v0 = b0 ? x[j0] : -INFINITY, v1 = b1 ? x[j1] : -INFINITY;
    const float t = fminf(fmaxf(targets[row], vmin), vmax);
    const float c0 = j0 <= NL ? ndtr_f32((e0 - t) / sigma) : 0.f;
    const float c1 = j1 <= NL ? ndtr_f32((e1 - t) / sigma) : 0.f;
    const float c2 = __shfl_down_sync(0xffffffffu, c0, 1);
    const float cdf_lo = __shfl_sync(0xffffffffu, c0, 0);
    const float cdf_hi = __shfl_sync(0xffffffffu, (NL & 1) ? c1 : c0, NL >> 1);
    const float rz = 1.0f / (cdf_hi - cdf_lo);                        // one reciprocal per row (all lanes hold it)
    const float p0 = b0 ? (c1 - c0) * rz : 0.f, p1 = b1 ? (c2 - c1) * rz : 0.f;
    const float m = warp_max(fmaxf(v0, v1));
    const float d0 = v0 - m, d1 = v1 - m;
    const float ex0 = b0 ? expf(d0) : 0.f, ex1 = b1 ? expf(d1) : 0.f;
    // three independent butterflies in one pass: sum exp, sum p, sum p * (v - m);
    // loss row = -sum p * ((v - m) - lse) = lse * sum p - sum p * (v - m)
    float s = ex0 + ex1, psum = p0 + p1, pd = (b0 ? p0 * d0 : 0.f) + (b1 ? p1 * d1 : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      psum += __shfl_xor_sync(0xffffffffu, psum, o);
      pd += __shfl_xor_sync(0xffffffffu, pd, o);
    }
    const float lse = logf(s);
    const float rs = 1.0f / s;
    const float g0 = gscale * (ex0 * rs * psum - p0), g1 = gscale * (ex1 * rs * psum - p1);   // softmax = exp(v - m) / s
    if (dlogits) {
      if (b0) dlogits[row * NL + j0] = g0;
      if (b1) dlogits[row * NL + j1] = g1;
    }
    if (dl_bf16) {
      if ((ld_bf16 & 1) == 0) {                                       // j0 even: one 4-byte store per lane, zero padded
        if (j0 < ld_bf16)
          *reinterpret_cast<uint32_t*>(dl_bf16 + row * ld_bf16 + j0) = pack_bf16(b0 ? g0 : 0.f, b1 ? g1 : 0.f);
      } else {
        if (j0 < ld_bf16) dl_bf16[row * ld_bf16 + j0] = __float2bfloat16_rn(b0 ? g0 : 0.f);
        if (j1 < ld_bf16) dl_bf16[row * ld_bf16 + j1] = __float2bfloat16_rn(b1 ? g1 : 0.f);
      }
    }
    if (lane == 0) acc += (double)(lse * psum - pd);
  }
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < HLG_WARPS; ++w) a += s_part[w];
    partials[blockIdx.x] = a;
  }
}

__global__ void hlgauss_loss_finalize_kernel(const double* __restrict__ partials, int nparts, double inv_n,
                                             float* __restrict__ loss) {
  __shared__ double s[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) a += partials[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(s[0] * inv_n);
}

static int hlg_grid(long long n) {
  long long g = (n + HLG_WARPS - 1) / HLG_WARPS;
  long long cap = (long long)num_sms() * 8;
  return (int)(g < cap ? g : cap);
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_hlgauss_value(const float* logits, long long n, int n_logits, const float* centers, float* value,
                                 mpx_stream_t stream) {
  MPX_REQUIRE(n > 0, "mpx_hlgauss_value: empty input");
  MPX_REQUIRE(n_logits > 0 && n_logits <= HLG_MAX_PER_LANE * 32, "mpx_hlgauss_value: n_logits=%d not in [1,128]",
              n_logits);
  hlgauss_value_kernel<<<hlg_grid(n), HLG_WARPS * 32, 0, (cudaStream_t)stream>>>(logits, n, n_logits, centers, value);
  return check_launch("mpx_hlgauss_value");
}

extern "C" size_t mpx_hlgauss_loss_workspace_bytes(long long n) { return (size_t)hlg_grid(n > 0 ? n : 1) * sizeof(double); }

extern "C" int mpx_hlgauss_loss(const float* logits, const float* targets, long long n, int n_logits,
                                const float* supports, float vmin, float vmax, float sigma, float grad_scale,
                                float* loss, float* dlogits, mpx_bf16* dlogits_bf16, int ld_bf16, void* workspace,
                                mpx_stream_t stream) {
  MPX_REQUIRE(!dlogits_bf16 || (ld_bf16 >= n_logits && ld_bf16 <= HLG_MAX_PER_LANE * 32),
              "mpx_hlgauss_loss: ld_bf16=%d must be in [n_logits,128]", ld_bf16);
  MPX_REQUIRE(n > 0, "mpx_hlgauss_loss: empty input");
  MPX_REQUIRE(n_logits > 0 && n_logits <= HLG_MAX_PER_LANE * 32, "mpx_hlgauss_loss: n_logits=%d not in [1,128]",
              n_logits);
  MPX_REQUIRE(sigma > 0.f && vmax > vmin, "mpx_hlgauss_loss: bad sigma/min/max");
  MPX_REQUIRE(loss && workspace, "mpx_hlgauss_loss: null pointer");
  const int grid = hlg_grid(n);
  double* partials = reinterpret_cast<double*>(workspace);
  if (n_logits < 64 && ld_bf16 <= 64)      // bin edge NL must live in lane NL >> 1 < 32
    hlgauss_loss2_kernel<<<grid, HLG_WARPS * 32, 0, (cudaStream_t)stream>>>(
        logits, targets, n, n_logits, supports, vmin, vmax, sigma, grad_scale / (float)n, dlogits,
        reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ld_bf16, partials);
  else
    hlgauss_loss_kernel<<<grid, HLG_WARPS * 32, 0, (cudaStream_t)stream>>>(
        logits, targets, n, n_logits, supports, vmin, vmax, sigma, grad_scale / (float)n, dlogits,
        reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ld_bf16, partials);
  hlgauss_loss_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, 1.0 / (double)n, loss);
  return check_launch("mpx_hlgauss_loss");
}
