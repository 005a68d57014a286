// Optimizer step and weight staging on flat fp32 parameter / gradient buffers.
//   AdamW + post-optimizer global-norm clip, exactly the optax chain of the reference
//   (mapox_trainer/optimizer.py:12-23: optax.adamw(linear_schedule(lr,0,N), b1,b2,eps,weight_decay) then
//   optax.clip_by_global_norm(max_norm) applied to the UPDATE), with the step counter kept on the device so the
//   whole minibatch step is CUDA-graph replayable.
//   prep_weights: fp32 master weights -> bf16 operand copies (optionally transposed / re-tiled) for the GEMMs.
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

__global__ void __launch_bounds__(256) adamw_update_kernel(
    const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu, float* __restrict__ nu,
    float* __restrict__ upd, long long n, const int* __restrict__ step, float lr0, float total_steps, float b1, float b2,
    float eps, float wd, float grad_scale, double* __restrict__ partials) {
  __shared__ double s_sum[256];
  const int t = step[0];
  const float frac = 1.0f - fminf((float)t, total_steps) / total_steps;
  const float lr = lr0 * frac;
  const float bc1 = 1.0f - powf(b1, (float)(t + 1));
  const float bc2 = 1.0f - powf(b2, (float)(t + 1));
  double acc = 0.0;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float gi = g[i] * grad_scale;
    const float m = b1 * mu[i] + (1.0f - b1) * gi;
    const float v = b2 * nu[i] + (1.0f - b2) * gi * gi;
    mu[i] = m;
    nu[i] = v;
    const float u = -lr * ((m / bc1) / (sqrtf(v / bc2) + eps) + wd * p[i]);
    upd[i] = u;
    acc += (double)u * (double)u;
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = s_sum[0];
}

// p += upd * min(1, max_norm / ||upd||) ; step += 1.  Each block re-reduces the partials in a fixed order.
__global__ void __launch_bounds__(256) apply_update_kernel(float* __restrict__ p, const float* __restrict__ upd,
                                                           long long n, const double* __restrict__ partials, int nparts,
                                                           float max_norm, int* __restrict__ step,
                                                           float* __restrict__ norm_out) {
  __shared__ double s_sum[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 256) a += partials[i];
  s_sum[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  const float norm = (float)sqrt(s_sum[0]);
  const float scale = (max_norm > 0.f && norm >= max_norm) ? max_norm / norm : 1.0f;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] += upd[i] * scale;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (norm_out) norm_out[0] = norm;
  }
}
__global__ void step_inc_kernel(int* step) { step[0] += 1; }

// ------------------------------------------------------------------------------------------------ weight staging
__global__ void prep_weights_kernel(const mpx_prep_desc* __restrict__ descs, int ndesc) {
  __shared__ float tile[32][33];
  const int d = blockIdx.y;
  if (d >= ndesc) return;
  const mpx_prep_desc ds = descs[d];
  const int tr = (ds.rows + 31) / 32, tc = (ds.cols + 31) / 32;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ds.dst);
  for (int tIdx = blockIdx.x; tIdx < tr * tc; tIdx += gridDim.x) {
    const int r0 = (tIdx / tc) * 32, c0 = (tIdx % tc) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
      const int r = r0 + rr, c = c0 + tx;
      tile[rr][tx] = (r < ds.rows && c < ds.cols) ? ds.src[(long long)r * ds.src_ld + c] : 0.f;
    }
    __syncthreads();
    // mode bit 0: transpose; bit 1: store the bf16 residual  bf16(w - f32(bf16(w)))  (lo half of a hi/lo split)
    const bool lo = (ds.transpose & 2) != 0;
    if (ds.transpose & 1) {
      for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + tx;
        float w = tile[tx][cc];
        if (lo) w = w - bf16_round(w);
        if (r < ds.rows && c < ds.cols) dst[(long long)c * ds.dst_ld + r] = __float2bfloat16_rn(w);
      }
    } else {
      for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + tx;
        float w = tile[rr][tx];
        if (lo) w = w - bf16_round(w);
        if (r < ds.rows && c < ds.cols) dst[(long long)r * ds.dst_ld + c] = __float2bfloat16_rn(w);
      }
    }
  }
}

static int opt_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 4;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// Rollout buffers are written time-major (T,B,E) by the decode loop; training reads batch-major minibatch rows.
// dst[b][t][:] = src[t][perm ? perm[b] : b][:]  with E bytes per item (the reference's global row permutation,
// rollout.py:169-181, fused with the layout change).
__global__ void transpose_tb_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                    const int32_t* __restrict__ perm, int T, int B, int E) {
  const long long total = (long long)T * B * E;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int e = (int)(i % E);
    const long long bt = i / E;
    const int t = (int)(bt % T);
    const int b = (int)(bt / T);
    const int sb = perm ? perm[b] : b;
    dst[i] = src[((long long)t * B + sb) * E + e];
  }
}
__global__ void transpose_tb4_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                     const int32_t* __restrict__ perm, int T, int B, int E4) {
  const long long total = (long long)T * B * E4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int e = (int)(i % E4);
    const long long bt = i / E4;
    const int t = (int)(bt % T);
    const int b = (int)(bt / T);
    const int sb = perm ? perm[b] : b;
    dst[i] = src[((long long)t * B + sb) * E4 + e];
  }
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_transpose_tb(const void* src, void* dst, const int32_t* perm, int T, int B, int item_bytes,
                                mpx_stream_t stream) {
  MPX_REQUIRE(src && dst && T > 0 && B > 0 && item_bytes > 0, "mpx_transpose_tb: bad arguments");
  const long long total = (long long)T * B * item_bytes;
  const bool w4 = (item_bytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 4 == 0);
  const long long n = w4 ? total / 4 : total;
  long long g = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (w4)
    transpose_tb4_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(src),
                                                                  reinterpret_cast<uint32_t*>(dst), perm, T, B,
                                                                  item_bytes / 4);
  else
    transpose_tb_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(src),
                                                                 reinterpret_cast<uint8_t*>(dst), perm, T, B, item_bytes);
  return check_launch("mpx_transpose_tb");
}

extern "C" size_t mpx_adamw_workspace_bytes(long long n) { return (size_t)opt_grid(n > 0 ? n : 1) * sizeof(double); }

extern "C" int mpx_adamw_step(float* params, const float* grads, float* mu, float* nu, float* update_scratch,
                              long long n, int* step, float lr0, float total_steps, float b1, float b2, float eps,
                              float weight_decay, float max_norm, float grad_scale, float* update_norm_out,
                              void* workspace, mpx_stream_t stream) {
  MPX_REQUIRE(params && grads && mu && nu && update_scratch && step && workspace && n > 0, "mpx_adamw_step: bad arguments");
  MPX_REQUIRE(total_steps > 0.f, "mpx_adamw_step: total_steps must be positive");
  const int grid = opt_grid(n);
  double* partials = reinterpret_cast<double*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  adamw_update_kernel<<<grid, 256, 0, s>>>(params, grads, mu, nu, update_scratch, n, step, lr0, total_steps, b1, b2, eps,
                                           weight_decay, grad_scale, partials);
  apply_update_kernel<<<grid, 256, 0, s>>>(params, update_scratch, n, partials, grid, max_norm, step, update_norm_out);
  step_inc_kernel<<<1, 1, 0, s>>>(step);
  return check_launch("mpx_adamw_step");
}

extern "C" int mpx_prep_weights(const mpx_prep_desc* descs_device, int ndesc, mpx_stream_t stream) {
  MPX_REQUIRE(descs_device && ndesc > 0, "mpx_prep_weights: bad arguments");
  dim3 grid(64, ndesc);
  prep_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(descs_device, ndesc);
  return check_launch("mpx_prep_weights");
}
