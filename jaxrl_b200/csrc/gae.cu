// GAE reverse scan + global advantage normalisation.
// Restates Rollout.calculate_advantage (reference mapox_trainer/rollout.py:128-167):
//   acc_T = V_T ; acc_t = r_t + d_t * ((1-lambda) * V_{t+1} + lambda * acc_{t+1}),  d_t = next_terminated ? 0 : gamma
//   targets = acc ; adv = targets - V_t ; optional (adv - mean) / (std + 1e-8) over the whole (B,T) array.
// HBM-bound (17 B / agent-step, +8 B for the normalise pass); arithmetic is strictly sequential in t per row
// and uses un-fused fp32 mul/add so the result is bit-identical to a plain fp32 loop.
//
// Layout: inputs are batch-major (B,T) exactly as RolloutState stores them (rollout.py:46-91) - the three
// swapaxes of the reference (rollout.py:143-156) disappear.  A CTA owns 32 agent rows and walks time backwards
// in tiles of 64 steps: coalesced tile load -> smem (pitch 65, conflict-free) -> warp 0 scans (lane = row)
// -> coalesced tile store.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int GAE_ROWS = 32;
constexpr int GAE_TT = 64;
constexpr int GAE_PITCH = GAE_TT + 1;
constexpr int GAE_THREADS = 128;

__global__ void __launch_bounds__(GAE_THREADS) gae_scan_kernel(
    const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ next_term,
    int B, int T, float discount, float lam, float one_minus_lam, float* __restrict__ adv_out,
    float* __restrict__ tgt_out, double* __restrict__ partials) {
  __shared__ float s_r[GAE_ROWS * GAE_PITCH];    // rewards in, targets out
  __shared__ float s_v1[GAE_ROWS * GAE_PITCH];   // V_{t+1} in, advantages out
  __shared__ float s_v0[GAE_ROWS * GAE_PITCH];   // V_t
  __shared__ float s_d[GAE_ROWS * GAE_PITCH];    // discount (0 or gamma)

  const int row0 = blockIdx.x * GAE_ROWS;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool scanner = tid < 32;
  const int my_row = row0 + lane;

  float acc = 0.f;
  double sum = 0.0, sumsq = 0.0;
  if (scanner && my_row < B) acc = values[(size_t)my_row * (T + 1) + T];

  const int ntiles = (T + GAE_TT - 1) / GAE_TT;
  for (int tile = ntiles - 1; tile >= 0; --tile) {
    const int t0 = tile * GAE_TT;
    const int tw = min(GAE_TT, T - t0);
    for (int idx = tid; idx < GAE_ROWS * GAE_TT; idx += GAE_THREADS) {
      int r = idx / GAE_TT, c = idx % GAE_TT;
      int row = row0 + r;
      if (row < B && c < tw) {
        size_t o = (size_t)row * T + t0 + c;
        size_t ov = (size_t)row * (T + 1) + t0 + c;
        s_r[r * GAE_PITCH + c] = rewards[o];
        s_v0[r * GAE_PITCH + c] = values[ov];
        s_v1[r * GAE_PITCH + c] = values[ov + 1];
        s_d[r * GAE_PITCH + c] = next_term[o] ? 0.0f : discount;
      }
    }
    __syncthreads();
    if (scanner && my_row < B) {
      for (int c = tw - 1; c >= 0; --c) {
        int i = lane * GAE_PITCH + c;
        float inner = __fadd_rn(__fmul_rn(one_minus_lam, s_v1[i]), __fmul_rn(lam, acc));
        acc = __fadd_rn(s_r[i], __fmul_rn(s_d[i], inner));
        float a = __fsub_rn(acc, s_v0[i]);
        s_r[i] = acc;
        s_v1[i] = a;
        sum += (double)a;
        sumsq += (double)a * (double)a;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < GAE_ROWS * GAE_TT; idx += GAE_THREADS) {
      int r = idx / GAE_TT, c = idx % GAE_TT;
      int row = row0 + r;
      if (row < B && c < tw) {
        size_t o = (size_t)row * T + t0 + c;
        tgt_out[o] = s_r[r * GAE_PITCH + c];
        adv_out[o] = s_v1[r * GAE_PITCH + c];
      }
    }
    __syncthreads();
  }
  if (scanner) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
    }
    if (lane == 0) {
      partials[2 * blockIdx.x] = sum;
      partials[2 * blockIdx.x + 1] = sumsq;
    }
  }
}

// Fixed-order reduction of the per-CTA partials -> stats[0] = sum(adv), stats[1] = sum(adv^2). Deterministic.
__global__ void gae_stats_kernel(const double* __restrict__ partials, int n, double* __restrict__ stats) {
  __shared__ double s0[256], s1[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    a += partials[2 * i];
    b += partials[2 * i + 1];
  }
  s0[threadIdx.x] = a;
  s1[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s0[threadIdx.x] += s0[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = s0[0];
    stats[1] = s1[0];
  }
}

__global__ void adv_normalize_kernel(float* __restrict__ adv, long long n, const double* __restrict__ stats,
                                     double count) {
  const double mean_d = stats[0] / count;
  double var = stats[1] / count - mean_d * mean_d;
  if (var < 0.0) var = 0.0;
  const float mean = (float)mean_d;
  const float denom = (float)sqrt(var) + 1e-8f;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<float4*>(adv + i);
    v.x = (v.x - mean) / denom;
    v.y = (v.y - mean) / denom;
    v.z = (v.z - mean) / denom;
    v.w = (v.w - mean) / denom;
    *reinterpret_cast<float4*>(adv + i) = v;
  }
  if (i < n) {   // ragged tail (n % 4 != 0): exactly one thread lands here
    for (long long j = i; j < n; ++j) adv[j] = (adv[j] - mean) / denom;
  }
}

}  // namespace mpx

using namespace mpx;

extern "C" size_t mpx_gae_workspace_bytes(int B, int T) {
  (void)T;
  return (size_t)(ceil_div(B, GAE_ROWS) * 2 + 2) * sizeof(double);
}

extern "C" int mpx_gae(const float* rewards, const float* values, const uint8_t* next_terminated, int B, int T,
                       double discount, double gae_lambda, float* advantages, float* targets, double* stats,
                       void* workspace, mpx_stream_t stream) {
  MPX_REQUIRE(B > 0 && T > 0, "mpx_gae: empty input B=%d T=%d", B, T);
  MPX_REQUIRE(rewards && values && next_terminated && advantages && targets && stats && workspace,
              "mpx_gae: null pointer");
  const int grid = ceil_div(B, GAE_ROWS);
  double* partials = reinterpret_cast<double*>(workspace);
  // (1 - gae_lambda) is folded in double before the scan, as the python-float expression in rollout.py:138 is.
  // discount / gae_lambda arrive as doubles (python floats in the reference) and are rounded to fp32 once each.
  const float one_minus = (float)(1.0 - gae_lambda);
  gae_scan_kernel<<<grid, GAE_THREADS, 0, (cudaStream_t)stream>>>(rewards, values, next_terminated, B, T,
                                                                  (float)discount, (float)gae_lambda, one_minus,
                                                                  advantages, targets, partials);
  gae_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, stats);
  return check_launch("mpx_gae");
}

extern "C" int mpx_adv_normalize(float* advantages, long long n, const double* stats, double count,
                                 mpx_stream_t stream) {
  MPX_REQUIRE(n > 0 && count > 0, "mpx_adv_normalize: empty input");
  MPX_REQUIRE((reinterpret_cast<uintptr_t>(advantages) & 15) == 0, "mpx_adv_normalize: advantages must be 16 B aligned");
  int grid = (int)((n / 4 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  adv_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(advantages, n, stats, count);
  return check_launch("mpx_adv_normalize");
}
