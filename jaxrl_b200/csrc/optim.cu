// Optimizer step and weight staging on flat fp32 parameter / gradient buffers.
//   AdamW + post-optimizer global-norm clip, exactly the optax chain of the reference
//   (mapox_trainer/optimizer.py:12-23: optax.adamw(linear_schedule(lr,0,N), b1,b2,eps,weight_decay) then
//   optax.clip_by_global_norm(max_norm) applied to the UPDATE), with the step counter kept on the device so the
//   whole minibatch step is CUDA-graph replayable.
//   prep_weights: fp32 master weights -> bf16 operand copies (optionally transposed / re-tiled) for the GEMMs.
#include "common.cuh"
#include "muon.h"
#include "ptx.cuh"

namespace mpx {

__global__ void __launch_bounds__(256) adamw_update_kernel(
    const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu, float* __restrict__ nu,
    float* __restrict__ upd, long long n, const int* __restrict__ step, float lr0, float total_steps, float b1, float b2,
    float eps, float wd, float grad_scale, int nesterov, double* __restrict__ partials) {
  __shared__ double s_sum[256];
  const int t = step[0];
  const float frac = 1.0f - fminf((float)t, total_steps) / total_steps;
  const float lr = lr0 * frac;
  const float bc1 = 1.0f - powf(b1, (float)(t + 1));
  const float bc1n = 1.0f - powf(b1, (float)(t + 2));
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
    // optax.scale_by_adam(nesterov=True): mu_hat = b1 * bias_corr(mu, t+2) + (1-b1) * bias_corr(g, t+1)
    const float mhat = nesterov ? (b1 * (m / bc1n) + (1.0f - b1) * (gi / bc1)) : (m / bc1);
    const float u = -lr * (mhat / (sqrtf(v / bc2) + eps) + wd * p[i]);
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
    step[0] += 1;          // nothing in this kernel reads the counter; the update kernels before it already have
  }
}

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

// Items of even size that are not a multiple of 4 bytes (the 11x11xK uint8 observation: 242 B) took the byte kernel
// above - one byte per thread-iteration, three 64-bit divisions each, 1.8 ms for the 507 MB observation buffer.
// Here a CTA assembles TB_TT consecutive time steps of ONE output row in shared memory - each warp fetches whole
// items (contiguous in the time-major source) as 16-bit words, two items in flight per warp - and writes the
// contiguous TB_TT * E byte segment of the batch-major row with 16-byte stores.
constexpr int TB_TT = 128;
__global__ void __launch_bounds__(256) transpose_tb_rows_kernel(const uint16_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                const int32_t* __restrict__ perm, int T, int B, int E2) {
  extern __shared__ __align__(16) uint16_t tb_sm[];          // TB_TT * E2
  const int b = blockIdx.x, t0 = blockIdx.y * TB_TT;
  const int nt = min(TB_TT, T - t0);
  const int sb = perm ? perm[b] : b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = warp; it < nt; it += 16) {
    const int it1 = it + 8;
    const uint16_t* s0 = src + ((long long)(t0 + it) * B + sb) * E2;
    const uint16_t* s1 = src + ((long long)(t0 + (it1 < nt ? it1 : it)) * B + sb) * E2;
    uint16_t r0[8], r1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = lane + 32 * k;
      r0[k] = e < E2 ? s0[e] : (uint16_t)0;
      r1[k] = e < E2 ? s1[e] : (uint16_t)0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int e = lane + 32 * k;
      if (e < E2) {
        tb_sm[it * E2 + e] = r0[k];
        if (it1 < nt) tb_sm[it1 * E2 + e] = r1[k];
      }
    }
  }
  __syncthreads();
  uint8_t* o = dst + ((long long)b * T + t0) * (2LL * E2);
  const int n16 = nt * E2;                                    // 16-bit words in the segment
  const int nvec = n16 >> 3;
  for (int i = threadIdx.x; i < nvec; i += 256) reinterpret_cast<uint4*>(o)[i] = reinterpret_cast<const uint4*>(tb_sm)[i];
  for (int i = (nvec << 3) + threadIdx.x; i < n16; i += 256) reinterpret_cast<uint16_t*>(o)[i] = tb_sm[i];
}

// ------------------------------------------------------------------------------------------------ Muon
// optax.contrib.muon as chained by the reference (optimizer.py:25-37): matrices labelled "muon" get
//   mu = beta*mu + (1-beta)*g ; mu_hat = beta*bc(mu, t+2) + (1-beta)*bc(g, t+1)           (nesterov)
//   X = mu_hat oriented rows <= cols ; X /= (||X||_F + eps) ; 5 x Newton-Schulz: A = X X^T, B = c1 A + c2 A A,
//   X = c0 X + B X ; update = -lr * ( sqrt(max(1, out/in)) * X + weight_decay * p )
// everything else goes through AdamW(nesterov=True); clip_by_global_norm then acts on the combined update.
// Newton-Schulz runs on fp32 matrices (as optax does).  Two implementations of its contractions: the tensor-core one with
// bf16 hi/lo operand pairs (muon_ns_tc.cu) and the batched SIMT fp32 GEMM below.  The chain is 15 dependent launches; with a
// handful of 128-row matrices (return_2_layers: 8) it is pure latency and the lighter SIMT CTAs finish sooner (244 vs 306 us
// per step, and they do not take 129 KB of shared memory from the kernels they overlap with), from a few dozen matrices on
// the tensor cores win (blog_32_layers 1314 -> 756 us, multitask 2418 -> 1176 us).  mpx_debug_set(5, v): 0 = by table
// size, 1 = SIMT, 2 = tensor cores.
int g_muon_ns_impl = 0;

__global__ void __launch_bounds__(256) muon_prepare_kernel(const MuonMat* __restrict__ mats, const float* __restrict__ g,
                                                           float* __restrict__ mu, float* __restrict__ X,
                                                           float* __restrict__ norm2, const int* __restrict__ step,
                                                           float beta, float grad_scale) {
  __shared__ float s_sum[256];
  const MuonMat mt = mats[blockIdx.y];
  const int t = step[0];
  const float bc1 = 1.0f - powf(beta, (float)(t + 1));
  const float bc2 = 1.0f - powf(beta, (float)(t + 2));
  const unsigned n = (unsigned)(mt.rows * mt.cols);
  float acc = 0.f;
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float gi = g[mt.off + i] * grad_scale;
    const float m = beta * mu[mt.off + i] + (1.0f - beta) * gi;
    mu[mt.off + i] = m;
    const float mh = beta * (m / bc2) + (1.0f - beta) * (gi / bc1);
    const int pr = (int)(i / (unsigned)mt.cols), pc = (int)(i % (unsigned)mt.cols);
    const long long xi = mt.transposed ? (long long)pc * mt.c + pr : (long long)i;
    X[mt.xoff + xi] = mh;
    acc += mh * mh;
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(norm2 + blockIdx.y, s_sum[0]);
}

// Batched Newton-Schulz GEMMs on row-major fp32 matrices; 64x64 tile, 16-wide k slab, 4x4 per thread, float4 smem reads.
// mode 0: A(r,r) += X X^T * inv_norm^2 over a 128-wide K split (A zeroed by the host; split-K keeps all SMs busy;
//         the first iteration folds the Frobenius normalisation in) ;
// mode 1: B(r,r) = c2 * A A + c1 * A ; mode 2: X'(r,c) = (B X + c0 * X) * s  (s = inv_norm on the first iteration)
constexpr int NS_KSPLIT = 128;
__global__ void __launch_bounds__(256) muon_ns_gemm_kernel(const MuonMat* __restrict__ mats, const float* __restrict__ Xin,
                                                           float* __restrict__ Xout, float* __restrict__ Abuf,
                                                           float* __restrict__ Bbuf, const float* __restrict__ norm2,
                                                           int mode, int first, int ntiles_n, float c0, float c1, float c2,
                                                           float eps) {
  __shared__ __align__(16) float sA[16][68], sB[16][68];
  const MuonMat mt = mats[blockIdx.z];
  const int r = mt.r, c = mt.c;
  const int M = r, N = (mode == 2) ? c : r, K = (mode == 0) ? c : r;
  const int nt = blockIdx.x % ntiles_n, ks = blockIdx.x / ntiles_n;     // ks > 0 only in mode 0
  const int m0 = blockIdx.y * 64, n0 = nt * 64;
  const int kbeg = (mode == 0) ? ks * NS_KSPLIT : 0;
  const int kend = (mode == 0) ? min(K, kbeg + NS_KSPLIT) : K;
  if (m0 >= M || n0 >= N || kbeg >= K) return;
  const float* A;  long long lda;      // A[i][k]
  const float* B;  long long sbk, sbj; // B[k][j] = B[k*sbk + j*sbj]
  const float* X = Xin + mt.xoff;
  if (mode == 0) { A = X; lda = c; B = X; sbk = 1; sbj = c; }
  else if (mode == 1) { A = Abuf + mt.aoff; lda = r; B = Abuf + mt.aoff; sbk = r; sbj = 1; }
  else { A = Bbuf + mt.aoff; lda = r; B = X; sbk = c; sbj = 1; }
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // software pipeline: the next k slab's global loads are in flight (registers) while the current slab is multiplied
  float pa[4], pb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      const int kk = i % 16, mm = i / 16;
      pa[q] = (m0 + mm < M && k0 + kk < kend) ? A[(long long)(m0 + mm) * lda + k0 + kk] : 0.f;
      const int kb = (sbj == 1) ? i / 64 : i % 16, nb = (sbj == 1) ? i % 64 : i / 16;
      pb[q] = (n0 + nb < N && k0 + kb < kend) ? B[(long long)(k0 + kb) * sbk + (long long)(n0 + nb) * sbj] : 0.f;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      sA[i % 16][i / 16] = pa[q];
      if (sbj == 1) sB[i / 64][i % 64] = pb[q];
      else sB[i % 16][i / 16] = pb[q];
    }
    __syncthreads();
    if (k0 + 16 < kend) fetch(k0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float inv = first ? 1.0f / (sqrtf(norm2[blockIdx.z]) + eps) : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= N) continue;
      if (mode == 0) atomicAdd(&Abuf[mt.aoff + (long long)mm * r + nn], acc[i][j] * inv * inv);
      else if (mode == 1) Bbuf[mt.aoff + (long long)mm * r + nn] = c2 * acc[i][j] + c1 * A[(long long)mm * r + nn];
      else Xout[mt.xoff + (long long)mm * c + nn] = (acc[i][j] + c0 * X[(long long)mm * c + nn]) * inv;
    }
  }
}

__global__ void __launch_bounds__(256) muon_finish_kernel(const MuonMat* __restrict__ mats, const float* __restrict__ X,
                                                          const float* __restrict__ p, float* __restrict__ upd,
                                                          const int* __restrict__ step, float lr0, float total_steps,
                                                          float wd, double* __restrict__ partials) {
  __shared__ double s_sum[256];
  const MuonMat mt = mats[blockIdx.y];
  const int t = step[0];
  const float lr = lr0 * (1.0f - fminf((float)t, total_steps) / total_steps);
  const float shape_scale = sqrtf(fmaxf(1.0f, (float)mt.cols / (float)mt.rows));   // sqrt(max(1, fan_out / fan_in))
  const unsigned n = (unsigned)(mt.rows * mt.cols);
  double acc = 0.0;
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int pr = (int)(i / (unsigned)mt.cols), pc = (int)(i % (unsigned)mt.cols);
    const long long xi = mt.transposed ? (long long)pc * mt.c + pr : (long long)i;
    const float u = -lr * (shape_scale * X[mt.xoff + xi] + wd * p[mt.off + i]);
    upd[mt.off + i] = u;
    acc += (double)u * (double)u;
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = s_sum[0];
}

// partials[blockIdx.x] = sum of squares of this block's grid-stride share of x (fixed order: identical on every rank that
// holds the same data)
__global__ void __launch_bounds__(256) sumsq_partials_kernel(const float* __restrict__ x, long long n,
                                                             double* __restrict__ partials) {
  __shared__ double s_sum[256];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double v = (double)x[i];
    acc += v * v;
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = s_sum[0];
}

}  // namespace mpx

using namespace mpx;

extern "C" int mpx_transpose_tb(const void* src, void* dst, const int32_t* perm, int T, int B, int item_bytes,
                                mpx_stream_t stream) {
  MPX_REQUIRE(src && dst && T > 0 && B > 0 && item_bytes > 0, "mpx_transpose_tb: bad arguments");
  const long long total = (long long)T * B * item_bytes;
  const bool w4 = (item_bytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 4 == 0);
  // even item size, rows and segments 16-byte aligned, item <= 512 B: the shared-memory row assembler
  const bool rows = !w4 && (item_bytes % 2 == 0) && item_bytes <= 512 && T % 8 == 0 &&
                    ((long long)T * item_bytes) % 16 == 0 && (reinterpret_cast<uintptr_t>(dst) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(src) % 2 == 0);
  if (rows) {
    const size_t smem = (size_t)TB_TT * item_bytes;
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
      cudaError_t e = cudaFuncSetAttribute(transpose_tb_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(transpose_tb_rows): %s", cudaGetErrorString(e));
        return MPX_ERR_CUDA;
      }
      attr_smem = smem;
    }
    transpose_tb_rows_kernel<<<dim3(B, ceil_div(T, TB_TT)), 256, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint16_t*>(src), reinterpret_cast<uint8_t*>(dst), perm, T, B, item_bytes / 2);
    return check_launch("mpx_transpose_tb");
  }
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
                                           weight_decay, grad_scale, 0, partials);
  apply_update_kernel<<<grid, 256, 0, s>>>(params, update_scratch, n, partials, grid, max_norm, step, update_norm_out);
  return check_launch("mpx_adamw_step");
}

extern "C" int mpx_prep_weights(const mpx_prep_desc* descs_device, int ndesc, mpx_stream_t stream) {
  MPX_REQUIRE(descs_device && ndesc > 0, "mpx_prep_weights: bad arguments");
  dim3 grid(64, ndesc);
  prep_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(descs_device, ndesc);
  return check_launch("mpx_prep_weights");
}

// Muon + AdamW(nesterov) partition (optimizer.py:25-37).  Flat buffers of n elements: [0, n_adam) are the AdamW
// parameters, the matrices described by `mats` (device array of nmat entries, see MuonMat) live in [n_adam, n).
// ns_ws: fp32 workspace >= 2*sum(r*c) + 2*sum(r*r) + nmat floats.  workspace >= mpx_muon_workspace_bytes(n, nmat).
extern "C" size_t mpx_muon_workspace_bytes(long long n, int nmat) {
  return (size_t)(opt_grid(n > 0 ? n : 1) + (size_t)nmat * 32) * sizeof(double);
}

// phases: bit 0 = the Muon-labelled matrices (momentum, Newton-Schulz, update + its squared-norm partials); bit 1 = the
// AdamW-labelled leaves, the post-update global-norm clip over BOTH partitions, the parameter write and the step
// counter.  A caller that has the matrices' gradients early (the layer weight gradients are complete long before the
// observation-encoder backward ends) runs phase 1 on a second stream and joins before phase 2.
extern "C" int mpx_muon_step_phased(float* params, const float* grads, float* mu, float* nu, float* update_scratch,
                                    long long n, long long n_adam, const void* mats, int nmat, long long x_total,
                                    long long a_total, int max_r, int max_c, float* ns_ws, int* step, float lr0,
                                    float total_steps, float adam_b1, float adam_b2, float eps, float weight_decay,
                                    float muon_beta, int ns_steps, float max_norm, float grad_scale,
                                    float* update_norm_out, void* workspace, int phases, mpx_stream_t stream) {
  MPX_REQUIRE(params && grads && mu && nu && update_scratch && step && workspace && n > 0, "mpx_muon_step: bad arguments");
  MPX_REQUIRE(n_adam >= 0 && n_adam <= n && (nmat == 0 || (mats && ns_ws)), "mpx_muon_step: bad partition");
  MPX_REQUIRE(total_steps > 0.f && ns_steps >= 1, "mpx_muon_step: bad schedule");
  MPX_REQUIRE(phases >= 1 && phases <= 3, "mpx_muon_step_phased: phases must be 1, 2 or 3");
  cudaStream_t s = (cudaStream_t)stream;
  double* partials = reinterpret_cast<double*>(workspace);
  const int agrid = n_adam > 0 ? opt_grid(n_adam) : 0;
  if (n_adam > 0 && (phases & 2))
    adamw_update_kernel<<<agrid, 256, 0, s>>>(params, grads, mu, nu, update_scratch, n_adam, step, lr0, total_steps,
                                              adam_b1, adam_b2, eps, weight_decay, grad_scale, 1, partials);
  int nparts = agrid + (nmat > 0 ? 32 * nmat : 0);
  if (nmat > 0 && (phases & 1)) {
    const MuonMat* mm = reinterpret_cast<const MuonMat*>(mats);
    float* X0 = ns_ws;
    float* X1 = X0 + x_total;
    float* Ab = X1 + x_total;
    float* Bb = Ab + a_total;
    float* norm2 = Bb + a_total;
    cudaMemsetAsync(norm2, 0, (size_t)nmat * sizeof(float), s);
    muon_prepare_kernel<<<dim3(32, nmat), 256, 0, s>>>(mm, grads, mu, X0, norm2, step, muon_beta, grad_scale);
    const float c0 = 3.4445f, c1 = -4.7750f, c2 = 2.0315f;
    float* xin = X0;
    float* xout = X1;
    for (int it = 0; it < ns_steps; ++it) {
      const int first = it == 0;
      if (g_muon_ns_impl == 2 || (g_muon_ns_impl == 0 && (nmat > 12 || max_r > 128))) {
        const int rc = muon_ns_tc_iteration(mm, nmat, max_r, max_c, xin, xout, Ab, Bb, norm2, first, c0, c1, c2, eps, s);
        if (rc) return rc;
      } else {
        const int tr = ceil_div(max_r, 64), tc = ceil_div(max_c, 64), nks = ceil_div(max_c, NS_KSPLIT);
        cudaMemsetAsync(Ab, 0, (size_t)a_total * sizeof(float), s);
        muon_ns_gemm_kernel<<<dim3(tr * nks, tr, nmat), 256, 0, s>>>(mm, xin, xout, Ab, Bb, norm2, 0, first, tr, c0, c1, c2, eps);
        muon_ns_gemm_kernel<<<dim3(tr, tr, nmat), 256, 0, s>>>(mm, xin, xout, Ab, Bb, norm2, 1, 0, tr, c0, c1, c2, eps);
        muon_ns_gemm_kernel<<<dim3(tc, tr, nmat), 256, 0, s>>>(mm, xin, xout, Ab, Bb, norm2, 2, first, tc, c0, c1, c2, eps);
      }
      float* tmp = xin; xin = xout; xout = tmp;
    }
    muon_finish_kernel<<<dim3(32, nmat), 256, 0, s>>>(mm, xin, params, update_scratch, step, lr0, total_steps,
                                                     weight_decay, partials + agrid);
  }
  if (phases & 2) {
    const int pgrid = opt_grid(n);
    apply_update_kernel<<<pgrid, 256, 0, s>>>(params, update_scratch, n, partials, nparts, max_norm, step, update_norm_out);
    }
  return check_launch("mpx_muon_step");
}

// Second half of the SHARDED Muon step (data-parallel ranks each own 1/world of the Muon-labelled matrices): phase 1
// (mpx_muon_step_phased, phases = 1, with the table of the LOCAL matrices) has written this rank's slice of
// update_scratch[n_adam:], the caller has all-gathered that partition; this call updates the AdamW-labelled leaves
// [0, n_adam), takes the global norm over the WHOLE update (every rank sums the same gathered data in the same order, so
// the clip factor is bit-identical everywhere), applies it and increments the step counter.
extern "C" int mpx_muon_apply_sharded(float* params, const float* grads, float* mu, float* nu, float* update_scratch,
                                      long long n, long long n_adam, int* step, float lr0, float total_steps, float adam_b1,
                                      float adam_b2, float eps, float weight_decay, float max_norm, float grad_scale,
                                      float* update_norm_out, void* workspace, mpx_stream_t stream) {
  MPX_REQUIRE(params && grads && mu && nu && update_scratch && step && workspace && n > 0 && n_adam >= 0 && n_adam <= n,
              "mpx_muon_apply_sharded: bad arguments");
  MPX_REQUIRE(total_steps > 0.f, "mpx_muon_apply_sharded: total_steps must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  double* partials = reinterpret_cast<double*>(workspace);
  const int agrid = n_adam > 0 ? opt_grid(n_adam) : 0;
  if (n_adam > 0)
    adamw_update_kernel<<<agrid, 256, 0, s>>>(params, grads, mu, nu, update_scratch, n_adam, step, lr0, total_steps,
                                              adam_b1, adam_b2, eps, weight_decay, grad_scale, 1, partials);
  int nparts = agrid;
  if (n > n_adam) {
    sumsq_partials_kernel<<<32, 256, 0, s>>>(update_scratch + n_adam, n - n_adam, partials + agrid);
    nparts += 32;
  }
  apply_update_kernel<<<opt_grid(n), 256, 0, s>>>(params, update_scratch, n, partials, nparts, max_norm, step, update_norm_out);
  return check_launch("mpx_muon_apply_sharded");
}

extern "C" int mpx_muon_step(float* params, const float* grads, float* mu, float* nu, float* update_scratch,
                             long long n, long long n_adam, const void* mats, int nmat, long long x_total,
                             long long a_total, int max_r, int max_c, float* ns_ws, int* step, float lr0,
                             float total_steps, float adam_b1, float adam_b2, float eps, float weight_decay,
                             float muon_beta, int ns_steps, float max_norm, float grad_scale, float* update_norm_out,
                             void* workspace, mpx_stream_t stream) {
  return mpx_muon_step_phased(params, grads, mu, nu, update_scratch, n, n_adam, mats, nmat, x_total, a_total, max_r, max_c,
                              ns_ws, step, lr0, total_steps, adam_b1, adam_b2, eps, weight_decay, muon_beta, ns_steps,
                              max_norm, grad_scale, update_norm_out, workspace, 3, stream);
}
